run() { N=$1; W=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 3 --workload $W "$@" > gpurun_out/r02d_bench_n${N}_$W.json 2> gpurun_out/r02d_bench_n${N}_$W.err; [ -s gpurun_out/r02d_bench_n${N}_$W.json ] || tail -5 gpurun_out/r02d_bench_n${N}_$W.err; }
run 8 p5_sah_4k --no-extras
run 8 p5_rgrid_4k --no-extras
run 8 p2_smallpt_64 --no-extras
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02d_bench_n8_*.json")):
    try:
        d=json.load(open(f)); e=d.get("e2e") or {}
        print(f.split("/")[-1], round(d["value"]), "Mrays/s", round(d["ms_per_step"],3), "ms | kernel max", round(d["kernel_ms_max_over_ranks"],3), "| verified", d["assembled_frame_verified"], "| e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),3), e.get("assembled_host_frame_verified"), e.get("phases_ms"), 'roof', (d.get('roofline') or {}).get('frac'))
    except Exception as ex: print(f, "ERR", ex)
PY
