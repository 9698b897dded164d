cd $GRAFT_REPO_ROOT
N="ncu --set full --clock-control none --import-source on"
python tools/profile_frame.py --workload p5_sah_4k --frames 3 > gpurun_out/plain_p5_sah_4k.log 2>&1 && $N -k regex:"k_whitted|k_montecarlo" -s 2 -c 1 -f -o gpurun_out/prof_p5_sah_4k python tools/profile_frame.py --workload p5_sah_4k --frames 3 > gpurun_out/ncu_p5_sah_4k.log 2>&1
python tools/profile_frame.py --workload p5_sah_4k --frames 3 --host-frame > gpurun_out/plain_p5_sah_4k_host.log 2>&1 && $N -k regex:"k_whitted|k_montecarlo" -s 2 -c 1 -f -o gpurun_out/prof_p5_sah_4k_host python tools/profile_frame.py --workload p5_sah_4k --frames 3 --host-frame > gpurun_out/ncu_p5_sah_4k_host.log 2>&1
python tools/profile_frame.py --workload p5_rgrid_4k --frames 3 > gpurun_out/plain_p5_rgrid_4k.log 2>&1 && $N -k regex:"k_whitted|k_montecarlo" -s 3 -c 2 -f -o gpurun_out/prof_p5_rgrid_4k python tools/profile_frame.py --workload p5_rgrid_4k --frames 3 > gpurun_out/ncu_p5_rgrid_4k.log 2>&1
python tools/profile_frame.py --workload p2_smallpt_64 --frames 3 > gpurun_out/plain_p2_smallpt_64.log 2>&1 && $N -k regex:"k_whitted|k_montecarlo" -s 2 -c 1 -f -o gpurun_out/prof_p2_smallpt_64 python tools/profile_frame.py --workload p2_smallpt_64 --frames 3 > gpurun_out/ncu_p2_smallpt_64.log 2>&1
python tools/profile_frame.py --workload p5_sah_4k --frames 4 --world 8 --col-block 32 > gpurun_out/plain_shard.log 2>&1 && $N -k regex:k_whitted -s 4 -c 2 -f -o gpurun_out/prof_shard8 python tools/profile_frame.py --workload p5_sah_4k --frames 4 --world 8 --col-block 32 > gpurun_out/ncu_shard.log 2>&1
python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/plain_launch.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/ncu_launch.log 2>&1
tail -n 3 gpurun_out/plain_*.log
ls -la gpurun_out/*.ncu-rep
