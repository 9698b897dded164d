import sys
sys.path.insert(0, '.')
import torch, rtb200
ctx = rtb200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
s = rtb200.PresetScene(2)
d = ctx.upload(s.flat)
w, h, spp = 640, 480, 32
buf = torch.empty((h, w, 3), dtype=torch.float32, device="cuda:0")
fr = rtb200.make_frame(w, h, samples=spp, seed=1)
for _ in range(3):
    r = d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)
print(r)
