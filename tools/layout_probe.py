"""rtb_render wall ms of one 4K frame by output layout and kind of host buffer (page-locked: the kernels store straight into
it; pageable: device frame + one copy).   python tools/layout_probe.py [--workload p5_sah_4k]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import rtb200  # noqa: E402
from bench import WORKLOADS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="p5_sah_4k")
ap.add_argument("--reps", type=int, default=8)
args = ap.parse_args()
wl = WORKLOADS[args.workload]
s = rtb200.PresetScene(wl["preset"], wl["algorithm"], wl["segments"])
W, H = wl["width"], wl["height"]
ctx = rtb200.Context(0)
d = ctx.upload(s.flat)
for name, layout in (("row-major", 0), ("reference (column-major)", rtb200.LAYOUT_REFERENCE), ("row-major rgb8", rtb200.OUTPUT_RGB8),
                     ("reference rgb8", rtb200.LAYOUT_REFERENCE | rtb200.OUTPUT_RGB8)):
    fr = rtb200.make_frame(W, H, samples=wl["samples"], layout=layout)
    shape = (W, H, 3) if layout & rtb200.LAYOUT_REFERENCE else (H, W, 3)
    pinned = rtb200.PinnedArray(((H * W * 3 + 3) // 4,) if layout & rtb200.OUTPUT_RGB8 else shape)
    host = pinned.array.view(np.uint8)[: H * W * 3].reshape(shape) if layout & rtb200.OUTPUT_RGB8 else pinned.array
    pageable = np.zeros(shape, np.uint8 if layout & rtb200.OUTPUT_RGB8 else np.float32)
    for kind, out in (("page-locked", host), ("pageable", pageable)):
        ms = []
        for _ in range(args.reps):
            t0 = time.perf_counter()
            _, st = d.render(s.camera, s.setting, fr, out=out)
            ms.append(((time.perf_counter() - t0) * 1e3, st["kernel_ms"]))
        ms = np.array(ms[3:])
        print(f"{args.workload} {name:28s} {kind:12s} call {ms[:,0].mean():6.2f} ms (kernels {ms[:,1].mean():.2f})", flush=True)
    assert np.array_equal(host, pageable), "page-locked and pageable frames differ"
    pinned.close()
d.close(); ctx.close(); s.close()
