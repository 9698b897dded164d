# the scaling line of one workload at N = 1, 2, 4, 8 on an 8-GPU box, plus other workloads at N = 8 (run under gpurun --gpus 8)
TAG=${1:-r02}
run() { # N workload extra...
  N=$1; W=$2; shift 2
  if [ "$N" = 1 ]; then python bench.py --gpus 1 --steps 20 --warmup 3 --workload $W "$@" > gpurun_out/scale_${TAG}_n${N}_$W.json 2> gpurun_out/scale_${TAG}_n${N}_$W.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N --steps 20 --warmup 3 --workload $W "$@" > gpurun_out/scale_${TAG}_n${N}_$W.json 2> gpurun_out/scale_${TAG}_n${N}_$W.err; fi
  [ -s gpurun_out/scale_${TAG}_n${N}_$W.json ] || tail -5 gpurun_out/scale_${TAG}_n${N}_$W.err
}
for N in 1 2 4 8; do run $N p5_sah_4k --no-extras; done
run 8 p5_rgrid_4k; run 8 p2_smallpt_64; run 8 p5_kd_4k; run 8 p5_fgrid_4k; run 4 p5_rgrid_4k; run 2 p5_rgrid_4k
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 8 --steps 20 --gather scatter > gpurun_out/scale_${TAG}_n8_scatter.json 2>/dev/null
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus 8 --steps 20 --gather nccl > gpurun_out/scale_${TAG}_n8_nccl.json 2>/dev/null
python - <<EOF
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_${TAG}_n*.json")):
    try:
        d=json.load(open(f))
        e=d.get("e2e") or {}
        print(f.split("/")[-1], round(d["value"]), "Mrays/s", round(d["ms_per_step"],3), "ms | kernel max", round(d["kernel_ms_max_over_ranks"],3), "| verified", d["assembled_frame_verified"], "| e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),3), e.get("assembled_host_frame_verified"), e.get("phases_ms"))
    except Exception as ex: print(f, "ERR", ex)
EOF
