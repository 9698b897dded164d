import sys, time
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
s = rtb200.PresetScene(5, "sah", 150)
fr = rtb200.make_frame(3840, 2880)
pin = rtb200.PinnedArray((2880, 3840, 3))
tpin = torch.empty((2880,3840,3), dtype=torch.float32).pin_memory().numpy()
page = np.zeros((2880, 3840, 3), np.float32)
for name, buf in [("rtb pinned", pin.array), ("torch pinned", tpin), ("pageable", page)]:
    for it in range(3):
        t0 = time.perf_counter(); d = ctx.upload(s.flat); t1 = time.perf_counter()
        _, st = d.render(s.camera, s.setting, fr, out=buf); t2 = time.perf_counter()
        d.close(); t3 = time.perf_counter()
        print(name, "upload %.2f render-call %.2f (kernel %.2f total-ev %.2f) free %.2f" % ((t1-t0)*1e3, (t2-t1)*1e3, st["kernel_ms"], st["total_ms"], (t3-t2)*1e3))
