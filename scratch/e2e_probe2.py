import sys, time, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
tag = os.environ.get("TAG", "")
for sc, (W, H) in [("5sah", (3840, 2880)), ("5sah", (400, 300)), ("5rgrid", (3840, 2880))]:
    s = rtb200.PresetScene(int(sc[0]), sc[1:], 150)
    for layout, name in [(0, "float"), (rtb200.OUTPUT_RGB8, "rgb8")]:
        fr = rtb200.make_frame(W, H, layout=layout)
        n = H * W * 3
        pin = rtb200.PinnedArray(((n * (4 if layout == 0 else 1) + 3) // 4,))
        buf = pin.array[: n].reshape(H, W, 3) if layout == 0 else pin.array.view(np.uint8)[: n].reshape(H, W, 3)
        ts = []
        for it in range(8):
            t0 = time.perf_counter(); d = ctx.upload(s.flat); t1 = time.perf_counter()
            img, st = d.render(s.camera, s.setting, fr, out=buf); t2 = time.perf_counter()
            d.close(); t3 = time.perf_counter()
            ts.append(((t3 - t0) * 1e3, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st["kernel_ms"], st["total_ms"]))
        best = min(ts[2:])
        print(tag, sc, "%dx%d" % (W, H), name, "step %.2f ms = upload %.2f + render call %.2f (kernel %.2f, events %.2f); checksum %.6g" % (best + (float(np.asarray(buf, np.float64).sum()),)), flush=True)
    s.close()
