import sys, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
W, H = 3840, 2880
buf = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
alg = os.environ.get("ALG", "sah")
s = rtb200.PresetScene(5, alg, 150)
d = ctx.upload(s.flat)
fr = rtb200.make_frame(W, H, rank=89, world=180, row_block=16, counters=int(os.environ.get("COUNTERS", "0")))
for i in range(3):
    r = d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)
    print(alg, "band rows 1424-1439:", r)
