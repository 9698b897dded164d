set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
RTB_UPLOAD_TIMING=1 timeout 120 python tools/upload_probe.py --reps 5 2>&1 | tail -4
timeout 200 python tools/multi_probe.py --devices 1,8 --same-device --reps 8 2>&1 | tee gpurun_out/multi_same2.log
timeout 600 python bench.py --no-extras > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -3 gpurun_out/c2_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/c2_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['value'], d['ms_per_step'], json.dumps(d['e2e'])[:900]); print(d['roofline']['frac'], d['roofline'].get('useful_lane_frac'), d.get('parity_checked'))
PY
