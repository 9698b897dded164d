import sys, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
W, H = 3840, 2880
buf = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
tag = os.environ.get("TAG", "")
def best(d, s, fr, n=4):
    ts = [d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)["kernel_ms"] for _ in range(n)]
    return min(ts[1:])
for preset, alg in [(5, "sah"), (5, "rgrid"), (5, "kd"), (4, "sah")]:
    s = rtb200.PresetScene(preset, alg, 150)
    d = ctx.upload(s.flat)
    full = best(d, s, rtb200.make_frame(W, H))
    band = best(d, s, rtb200.make_frame(W, H, rank=89, world=180, row_block=16))
    sh = [best(d, s, rtb200.make_frame(W, H, rank=r, world=8, row_block=16), 3) for r in range(8)]
    print(tag, preset, alg, "full %.2f band %.2f shard8 max %.2f mean %.2f" % (full, band, max(sh), sum(sh) / 8))
    d.close(); s.close()
