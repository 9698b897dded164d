import sys, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
W, H = 3840, 2880
buf = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
tag = os.environ.get("TAG", "")
RB = int(os.environ.get("RB", "8"))
def runs(d, s, fr, n=5):
    return [d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)["kernel_ms"] for _ in range(n)]
which = os.environ.get("SCENES", "5sah,5rgrid,5kd,4sah").split(",")
for sc in which:
    preset, alg = int(sc[0]), sc[1:]
    s = rtb200.PresetScene(preset, alg, 150)
    d = ctx.upload(s.flat)
    full = runs(d, s, rtb200.make_frame(W, H), 6)
    sh = [runs(d, s, rtb200.make_frame(W, H, rank=r, world=8, row_block=RB), 5) for r in range(8)]
    small = runs(d, s, rtb200.make_frame(400, 300), 6)
    print(tag, preset, alg, "full", " ".join("%.2f" % t for t in full[1:]), "| shard8 max %.2f mean %.2f (last runs: %s) | 400x300 %s" % (
        max(min(x[2:]) for x in sh), sum(min(x[2:]) for x in sh) / 8, " ".join("%.2f" % x[-1] for x in sh), " ".join("%.3f" % t for t in small[1:])))
    d.close(); s.close()
