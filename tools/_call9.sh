T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
timeout 400 $T bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > gpurun_out/c9_n8.json 2> gpurun_out/c9_n8.err || tail -5 gpurun_out/c9_n8.err
RTB_GROUP_STORE=0 timeout 400 $T bench.py --gpus 8 --steps 20 --warmup 3 --no-extras > gpurun_out/c9_n8_nogroup.json 2> gpurun_out/c9_n8_nogroup.err || tail -5 gpurun_out/c9_n8_nogroup.err
python - <<'PY'
import json
for f in ("gpurun_out/c9_n8.json", "gpurun_out/c9_n8_nogroup.json"):
    for l in open(f):
        if l.startswith('{'):
            d=json.loads(l); e=d['e2e']; print(f, 'value', round(d['value']), round(d['ms_per_step'],3), 'kernel', round(d['kernel_ms_max_over_ranks'],3), 'verified', d['assembled_frame_verified'], 'e2e', round(e['value']), round(e['ms_per_step'],3), e.get('assembled_host_frame_verified'), e.get('phases_ms'), 'roof', d['roofline'].get('frac'))
PY
