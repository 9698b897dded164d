import sys, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
tag = os.environ.get("TAG", "")
RB = int(os.environ.get("RB", "8"))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
buf = torch.empty((2880, 3840, 3), dtype=torch.float32, device="cuda:0")
def runs(d, s, fr, n):
    return [d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)["kernel_ms"] for _ in range(n)]
for sc in os.environ.get("SCENES", "5sah,5rgrid,5kd,5fgrid,4sah,4rgrid").split(","):
    preset, alg = int(sc[0]), sc[1:]
    s = rtb200.PresetScene(preset, alg, 150)
    d = ctx.upload(s.flat)
    for (W, H) in [(400, 300), (800, 600), (1280, 960)]:
        t = runs(d, s, rtb200.make_frame(W, H), 14)
        print(tag, sc, "%dx%d" % (W, H), " ".join("%.2f" % x for x in t), flush=True)
    t = runs(d, s, rtb200.make_frame(3840, 2880), 8)
    print(tag, sc, "4K full", " ".join("%.2f" % x for x in t), flush=True)
    sh = [runs(d, s, rtb200.make_frame(3840, 2880, rank=r, world=8, row_block=RB), 8) for r in range(8)]
    print(tag, sc, "4K shard8 per-rank (min of last 5 | last):", " ".join("%.2f|%.2f" % (min(x[3:]), x[-1]) for x in sh), "max %.2f" % max(min(x[3:]) for x in sh), flush=True)
    d.close(); s.close()
