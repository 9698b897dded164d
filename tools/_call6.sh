timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --no-extras > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; tail -3 gpurun_out/c6_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/c6_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('value', d['value'], d['ms_per_step'], 'e2e', e['value'], e['ms_per_step'], 'pageable', e['pageable_scene_arrays']['ms_per_step'], 'rgb8', e['rgb8_output_stage']['ms_per_step']); print(d['roofline']['frac'], d['roofline'].get('useful_lane_frac'), d.get('parity_checked'), d['first_frame_ms'], d['moving_camera_ms'])
PY
