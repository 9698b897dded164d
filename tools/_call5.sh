P="timeout 200 python tools/multi_probe.py --devices 1 --reps 10"
for wl in p5_sah_4k; do
for pm in 0 1; do for fd in 6 9 12 16; do RTB_GROUP_PERMUTE=$pm RTB_FLOOR_DELTA_GROUP=$fd $P --workload $wl 2>&1 | grep resident | sed "s/^/$wl permute=$pm floor=$fd /"; done; done
done | tee gpurun_out/group_permute_sweep.log
for wl in p5_rgrid_4k p5_kd_4k p5_fgrid_4k p4_sah_4k; do
RTB_GROUP_STORE=0 $P --workload $wl 2>&1 | grep resident | sed "s/^/$wl group=0 /"
for pm in 0 1; do RTB_GROUP_PERMUTE=$pm $P --workload $wl 2>&1 | grep resident | sed "s/^/$wl group=1 floor=9 permute=$pm /"; done
done | tee -a gpurun_out/group_permute_sweep.log
