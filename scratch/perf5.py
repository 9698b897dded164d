import sys, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
tag = os.environ.get("TAG", "")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
buf = torch.empty((2880, 3840, 3), dtype=torch.float32, device="cuda:0")
def runs(d, s, fr, n):
    return [d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)["kernel_ms"] for _ in range(n)]
sizes = [tuple(int(v) for v in z.split("x")) for z in os.environ.get("SIZES", "400x300").split(",")]
for sc in os.environ.get("SCENES", "5sah,5kd,5fgrid").split(","):
    preset, alg = int(sc[0]), sc[1:]
    s = rtb200.PresetScene(preset, alg, 150)
    d = ctx.upload(s.flat)
    for (W, H) in sizes:
        t = runs(d, s, rtb200.make_frame(W, H), 12)
        print(tag, sc, "%dx%d" % (W, H), " ".join("%.2f" % x for x in t), flush=True)
    d.close(); s.close()
