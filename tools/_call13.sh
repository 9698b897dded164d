timeout 900 python bench.py > gpurun_out/r02d_bench_n1.json 2> gpurun_out/r02d_bench_n1.err; tail -2 gpurun_out/r02d_bench_n1.err
timeout 600 python bench.py --workload p5_rgrid_4k --no-extras > gpurun_out/r02d_bench_n1_rgrid.json 2> gpurun_out/r02d_bench_n1_rgrid.err
timeout 600 python bench.py --workload p2_smallpt_64 --no-extras > gpurun_out/r02d_bench_n1_smallpt.json 2> gpurun_out/r02d_bench_n1_smallpt.err
timeout 600 python bench.py --workload pt_r2000_150 > gpurun_out/r02d_bench_pt.json 2> gpurun_out/r02d_bench_pt.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02d_bench_reference_arm.json 2> gpurun_out/r02d_bench_reference_arm.err
python - <<'PY'
import json
for f in ("r02d_bench_n1","r02d_bench_n1_rgrid","r02d_bench_n1_smallpt","r02d_bench_pt","r02d_bench_reference_arm"):
    try:
        for l in open(f"gpurun_out/{f}.json"):
            if l.startswith('{'):
                d=json.loads(l); e=d.get('e2e') or {}; r=d.get('roofline') or {}; c=d.get('cpu_baseline') or {}
                print(f, 'value', round(d['value'],1), d['unit'], round(d.get('ms_per_step',0),3), 'e2e', round(e.get('value',0),1), round(e.get('ms_per_step',0),3), 'roof', r.get('frac'), 'cpu', c.get('value'), c.get('cores'), 'parity', d.get('parity_checked'), 'first', d.get('first_frame_ms'))
    except Exception as ex: print(f, 'ERR', ex)
PY
