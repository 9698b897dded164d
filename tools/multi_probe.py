"""Phases of the end-to-end call on n devices driven by one host thread (rtb_multi_*): upload / render / free, ms.

  python tools/multi_probe.py [--workload p5_sah_4k] [--devices 1,2,4,8] [--reps 8]
"""
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
ap.add_argument("--devices", default="1,2")
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--same-device", action="store_true", help="n contexts on device 0 (one-GPU box: the host-side cost of the calls)")
args = ap.parse_args()
wl = WORKLOADS[args.workload]
s = rtb200.PresetScene(wl["preset"], wl["algorithm"], wl["segments"])
W, H = wl["width"], wl["height"]
pinned = rtb200.PinnedArray((H, W, 3))
for n in [int(v) for v in args.devices.split(",")]:
    m = rtb200.MultiContext(n, [0] * n) if args.same_device else rtb200.MultiContext(n)
    fr = rtb200.make_frame(W, H, samples=wl["samples"])
    rows = []
    for i in range(args.reps):
        t0 = time.perf_counter()
        d = m.upload(s.flat)
        t1 = time.perf_counter()
        _, st = d.render(s.camera, s.setting, fr, out=pinned.array)
        t2 = time.perf_counter()
        d.close()
        t3 = time.perf_counter()
        rows.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, st["kernel_ms"], st["total_ms"]))
    r = np.array(rows[3:])
    print(f"{args.workload} n={n}: upload {r[:,0].mean():.2f} render {r[:,1].mean():.2f} free {r[:,2].mean():.2f} ms | slowest device: kernels {r[:,3].mean():.2f} "
          f"whole call {r[:,4].mean():.2f} ms | step {r[:,:3].sum(axis=1).mean():.2f} ms", flush=True)
    # scene resident: render only
    d = m.upload(s.flat)
    t = []
    for i in range(args.reps):
        t0 = time.perf_counter()
        _, st = d.render(s.camera, s.setting, fr, out=pinned.array)
        t.append(((time.perf_counter() - t0) * 1e3, st["kernel_ms"]))
    t = np.array(t[3:])
    print(f"   scene resident: render {t[:,0].mean():.2f} ms (kernels {t[:,1].mean():.2f})", flush=True)
    d.close(); m.close()
