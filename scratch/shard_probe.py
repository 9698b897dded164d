import sys, time
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
s = rtb200.PresetScene(5, "sah", 150)
d = ctx.upload(s.flat)
W, H = 3840, 2880
buf = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
def run(world, rank, rb):
    fr = rtb200.make_frame(W, H, rank=rank, world=world, row_block=rb)
    best = None
    for _ in range(4):
        r = d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)
        best = r if best is None or r["kernel_ms"] < best["kernel_ms"] else best
    return best
for world, rb in [(1, 16), (2, 16), (2, 8), (2, 64), (2, 360), (4, 16), (8, 16), (8, 8), (8, 40)]:
    res = [run(world, r, rb) for r in range(world)]
    print("world", world, "rb", rb, "ms", [round(x["kernel_ms"], 2) for x in res], "Mrays", [round(x["n_rays"] / 1e6, 2) for x in res], "sum ms %.2f" % sum(x["kernel_ms"] for x in res))
