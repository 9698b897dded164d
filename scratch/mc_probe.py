import sys, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for preset, w, h, spp in [(2, 1280, 960, 64), (2, 400, 300, 256), (2, 3840, 2880, 16), (1, 1280, 960, 64), (3, 400, 300, 16), (3, 1280, 960, 8)]:
    s = rtb200.PresetScene(preset)
    d = ctx.upload(s.flat)
    buf = torch.empty((h, w, 3), dtype=torch.float32, device="cuda:0")
    fr = rtb200.make_frame(w, h, samples=spp, seed=1)
    best = None
    for _ in range(3):
        r = d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)
        best = r if best is None or r["kernel_ms"] < best["kernel_ms"] else best
    print("preset", preset, f"{w}x{h}x{spp}spp", "ms %.2f" % best["kernel_ms"], "Mrays/s %.0f" % (best["n_rays"] / best["kernel_ms"] / 1e3), "rays/sample %.2f" % (best["n_rays"] / (w * h * spp)))
    d.close(); s.close()
