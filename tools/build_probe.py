"""Accelerator build + upload time, host builders vs device builders (ms, median of reps)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rtb200
ctx = rtb200.Context(0)
def timed(fn, reps=5):
    t = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); t.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(t))
for preset, alg in ((5, "sah"), (4, "sah"), (5, "rgrid"), (5, "fgrid"), (4, "rgrid")):
    def host_path():
        s = rtb200.PresetScene(preset, alg, 150); d = ctx.upload(s.flat); d.close(); s.close()
    def device_path():
        (rtb200.set_kd_on_device if alg == "sah" else rtb200.set_grid_on_device)(True)
        try:
            s = rtb200.PresetScene(preset, alg, 150)
        finally:
            (rtb200.set_kd_on_device if alg == "sah" else rtb200.set_grid_on_device)(False)
        d = ctx.upload(s.flat); d.close(); s.close()
    s = rtb200.PresetScene(preset, alg, 150); hb = s.build_ms; s.close()
    print(f"preset {preset} {alg}: scene generation + host build + upload {timed(host_path):.1f} ms (host accelerator build alone {hb:.1f} ms) | generation + upload with the device build {timed(device_path):.1f} ms", flush=True)
