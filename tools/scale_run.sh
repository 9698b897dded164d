# bench.py at N ranks for a few workloads (run under gpurun --gpus N); writes gpurun_out/scale_<tag>_n<N>_<workload>.json
N=$1; TAG=${2:-r02}; shift 2 || true
WL=${*:-p5_sah_4k p5_rgrid_4k p2_smallpt_64}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for w in $WL; do
  $T bench.py --gpus $N --steps 20 --warmup 3 --workload $w > gpurun_out/scale_${TAG}_n${N}_$w.json 2> gpurun_out/scale_${TAG}_n${N}_$w.err || tail -5 gpurun_out/scale_${TAG}_n${N}_$w.err
done
python - <<EOF
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_${TAG}_n${N}_*.json")):
    try:
        d=json.load(open(f))
        e=d.get("e2e") or {}
        print(f.split("/")[-1], round(d["value"]), "Mrays/s", round(d["ms_per_step"],3), "ms | kernel max", round(d["kernel_ms_max_over_ranks"],3), "| verified", d["assembled_frame_verified"], "| e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),3), e.get("assembled_host_frame_verified"), e.get("phases_ms"))
    except Exception as ex: print(f, "ERR", ex)
EOF
