set -x
timeout 600 python -m pytest tests -m gpu -x -q -k "pretest or multi or upload or ipc or global_layout" 2>&1 | tail -5
timeout 300 bash tools/ab_probe.sh librtb200_cmp1.so 2>&1 | tee gpurun_out/ab_cmp.log
RTB_UPLOAD_TIMING=1 timeout 120 python tools/upload_probe.py --reps 6 2>&1 | tee gpurun_out/upload_laps.log
timeout 200 python tools/multi_probe.py --devices 1,2,4,8 --same-device --reps 8 2>&1 | tee gpurun_out/multi_same.log
