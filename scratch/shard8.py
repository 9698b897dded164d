# per-rank kernel ms of the 8-way row shards of the 4K frame, one after the other on one GPU (what bounds the N = 8 step)
import sys, os
sys.path.insert(0, '.')
import torch, rtb200
ctx = rtb200.Context(0)
tag = os.environ.get("TAG", "")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
buf = torch.empty((2880, 3840, 3), dtype=torch.float32, device="cuda:0")
def runs(d, s, fr, n):
    return [d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)["kernel_ms"] for _ in range(n)]
for sc in os.environ.get("SCENES", "5sah,5kd,4sah,5rgrid").split(","):
    s = rtb200.PresetScene(int(sc[0]), sc[1:], 150)
    d = ctx.upload(s.flat)
    world = 8
    for cb in [int(c) for c in os.environ.get("COLS", "0,32").split(",")]:
        sh = [runs(d, s, rtb200.make_frame(3840, 2880, rank=r, world=world, row_block=8, col_block=cb), 6) for r in range(world)]
        print(tag, sc, "world", world, "col_block", cb, "per-rank min of last 3:", " ".join("%.2f" % min(x[3:]) for x in sh), "max %.2f" % max(min(x[3:]) for x in sh), flush=True)
    d.close(); s.close()
