P="timeout 200 python tools/multi_probe.py --devices 1 --reps 10"
for fd in 9 0; do RTB_FLOOR_DELTA_HOST=$fd $P --workload p2_smallpt_64 2>&1 | grep resident | sed "s/^/p2_smallpt_64 floor_host=$fd /"; done
for wl in p5_rgrid_4k p5_kd_4k p5_fgrid_4k p4_sah_4k; do for rep in 1 2; do for fd in 9 0; do RTB_FLOOR_DELTA_GROUP=$fd $P --workload $wl 2>&1 | grep resident | sed "s/^/$wl floor_group=$fd /"; done; done; done
