set -x
timeout 900 python -m pytest tests/test_gpu_frame.py -m gpu -x -q 2>&1 | tail -5
for g in 0 1; do for wl in p5_sah_4k p5_rgrid_4k p5_kd_4k p5_fgrid_4k; do RTB_GROUP_STORE=$g timeout 200 python tools/multi_probe.py --workload $wl --devices 1 --reps 10 2>&1 | sed "s/^/group=$g /"; done; done | tee gpurun_out/group_store_ab.log
