for wl in p5_sah_4k p5_rgrid_4k; do
RTB_GROUP_STORE=0 timeout 200 python tools/multi_probe.py --workload $wl --devices 1 --reps 10 2>&1 | grep resident | sed "s/^/$wl group=0 /"
for fd in 0 3 6 9 12 16; do RTB_GROUP_STORE=1 RTB_FLOOR_DELTA_GROUP=$fd timeout 200 python tools/multi_probe.py --workload $wl --devices 1 --reps 10 2>&1 | grep resident | sed "s/^/$wl group=1 floor=$fd /"; done
done | tee gpurun_out/group_floor_sweep.log
