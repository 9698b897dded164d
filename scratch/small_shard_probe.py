# one 1/8 shard (rank 4: owns vanishing-point rows) of the 4K SAH frame, a few frames: target for ncu of the latency tiers
import sys
sys.path.insert(0, '.')
import torch, rtb200
ctx = rtb200.Context(0)
s = rtb200.PresetScene(5, "sah", 150)
d = ctx.upload(s.flat)
buf = torch.empty((2880, 3840, 3), dtype=torch.float32, device="cuda:0")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for (fr, n) in [(rtb200.make_frame(3840, 2880, rank=4, world=8, row_block=8), 6), (rtb200.make_frame(400, 300), 6)]:
    t = [d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)["kernel_ms"] for _ in range(n)]
    print(" ".join("%.3f" % x for x in t))
