import sys
sys.path.insert(0, '.')
import numpy as np, rtb200
ctx = rtb200.Context(0)
for alg in ["sah", "rgrid"]:
    s = rtb200.PresetScene(5, alg, 150)
    d = ctx.upload(s.flat)
    W, H = 3840, 2880
    img, st = d.render(s.camera, s.setting, rtb200.make_frame(W, H, counters=2))
    cyc, rays, work = img[..., 0], img[..., 1], img[..., 2]
    # per-warp (8x4 tile) max cycles
    wc = cyc.reshape(H // 4, 4, W // 8, 8).max(axis=(1, 3))
    ww = work.reshape(H // 4, 4, W // 8, 8)
    print(alg, "kernel_ms", st["kernel_ms"], "rays/px mean %.2f max %.0f" % (rays.mean(), rays.max()))
    print(" work/px mean %.0f p99 %.0f max %.0f" % (work.mean(), np.percentile(work, 99), work.max()))
    print(" warp cycles: mean %.0f p50 %.0f p99 %.0f p99.9 %.0f max %.0f  (1e6 cyc = 0.5 ms)" % (wc.mean(), np.percentile(wc, 50), np.percentile(wc, 99), np.percentile(wc, 99.9), wc.max()))
    iy, ix = np.unravel_index(np.argmax(wc), wc.shape)
    print(" slowest warp at tile row %d col %d (pixel y~%d x~%d); its lanes' work: max %.0f sum %.0f; rays max %.0f" % (iy, ix, iy * 4, ix * 8, ww[iy, :, ix, :].max(), ww[iy, :, ix, :].sum(), rays.reshape(H // 4, 4, W // 8, 8)[iy, :, ix, :].max()))
    # rows profile
    rowmax = wc.max(axis=1)
    top = np.argsort(rowmax)[-8:]
    print(" slowest tile rows:", [(int(t) * 4, int(rowmax[t])) for t in top])
    sw = ww.sum(axis=(1, 3)); mw = ww.max(axis=(1, 3))
    print(" SIMT: sum over warps of 32*max_lane_work / sum work = %.2f" % (32 * mw.sum() / sw.sum()))
    d.close(); s.close()
