timeout 900 python -m pytest tests/test_gpu_frame.py tests/test_gpu_parity.py -m gpu -x -q -k "global or shard or multi or ipc or column" 2>&1 | tail -3
P="python tools/frame_probe.py --world 8 --col-block 32 --reps 7 --global-frame"
for rep in 1 2; do for g in 0 1; do RTB_GROUP_STORE_PEER=$g $P --workloads p5_sah_4k,p5_rgrid_4k --tag "peer_group=$g" | sed 's/first-frame.*| steady/steady/'; done; done | tee gpurun_out/group_peer_1gpu.log
RTB_GROUP_STORE_PEER=1 python tools/frame_probe.py --world 2 --reps 7 --global-frame --workloads p5_sah_4k --tag "peer_group=1" | sed 's/first-frame.*| steady/steady/' | tee -a gpurun_out/group_peer_1gpu.log
RTB_GROUP_STORE_PEER=0 python tools/frame_probe.py --world 2 --reps 7 --global-frame --workloads p5_sah_4k --tag "peer_group=0" | sed 's/first-frame.*| steady/steady/' | tee -a gpurun_out/group_peer_1gpu.log
