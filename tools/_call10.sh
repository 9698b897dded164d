timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for wl in p5_sah_4k p5_rgrid_4k p5_kd_4k p5_fgrid_4k; do RTB_UPLOAD_TIMING=1 timeout 100 python tools/upload_probe.py --reps 6 --pin --workload $wl 2>&1 | tail -2; done
RTB_UPLOAD_TIMING=1 timeout 100 python tools/upload_probe.py --reps 6 2>&1 | tail -2
timeout 600 python bench.py --no-extras --no-cpu > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; tail -3 gpurun_out/c10_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/c10_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print('value', d['value'], d['ms_per_step'], 'e2e', e['value'], e['ms_per_step'], 'pageable', e['pageable_scene_arrays']['ms_per_step'], 'rgb8', e['rgb8_output_stage']['ms_per_step'])
PY
