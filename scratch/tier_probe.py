import sys, os
sys.path.insert(0, '.')
import numpy as np, rtb200
ctx = rtb200.Context(0)
W, H = [int(v) for v in os.environ.get("SIZE", "3840x2880").split("x")]
tag = os.environ.get("TAG", "")
for sc in os.environ.get("SCENES", "5sah,5rgrid").split(","):
    preset, alg = int(sc[0]), sc[1:]
    s = rtb200.PresetScene(preset, alg, 150)
    d = ctx.upload(s.flat)
    fr = rtb200.make_frame(W, H, counters=2)
    maps = []
    for k in range(4):
        img, st = d.render(s.camera, s.setting, fr)
        cyc = img[:, :, 0].astype(np.float64)
        th, tw = H // 4, W // 8
        tile = cyc[:th * 4, :tw * 8].reshape(th, 4, tw, 8).max(axis=(1, 3)).ravel()
        maps.append(tile)
        print(tag, sc, "frame", k, "kernel_ms %.3f" % st["kernel_ms"], "max tile Mcyc %.3f" % (tile.max() / 1e6), flush=True)
    a, b = maps[0], maps[3]
    order = np.argsort(-a)
    for lo, hi in [(0, 16), (16, 64), (64, 128), (128, 256), (256, 512), (512, 1024), (1024, 2048), (2048, 4096)]:
        sel = order[lo:hi]
        print(tag, sc, "tiles ranked %4d..%4d by raster-frame cost: raster Mcyc mean %.3f max %.3f | tiered frame Mcyc mean %.3f max %.3f | ratio %.2f" % (
            lo, hi, a[sel].mean() / 1e6, a[sel].max() / 1e6, b[sel].mean() / 1e6, b[sel].max() / 1e6, a[sel].mean() / b[sel].mean()))
    d.close(); s.close()
