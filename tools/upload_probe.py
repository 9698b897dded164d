"""Host-side cost of rtb_scene_upload (+ rtb_scene_free): wall ms per call, and the library's own laps on stderr.

  RTB_UPLOAD_TIMING=1 python tools/upload_probe.py [--workload p5_sah_4k] [--reps 8]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtb200  # noqa: E402
from bench import WORKLOADS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="p5_sah_4k")
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--pin", action="store_true", help="page-lock the scene arrays in place (rtb_flat_scene.arrays_page_locked)")
args = ap.parse_args()
wl = WORKLOADS[args.workload]
s = rtb200.PresetScene(wl["preset"], wl["algorithm"], wl["segments"])
ctx = rtb200.Context(0)
if args.pin:
    s.pin()
for i in range(args.reps):
    t0 = time.perf_counter()
    d = ctx.upload(s.flat)
    t1 = time.perf_counter()
    d.close()
    t2 = time.perf_counter()
    print(f"{args.workload} upload {1e3 * (t1 - t0):.3f} ms, free {1e3 * (t2 - t1):.3f} ms", flush=True)
ctx.close(); s.close()
