import sys, time, os
sys.path.insert(0, '.')
import numpy as np, rtb200
ctx = rtb200.Context(0)
s = rtb200.PresetScene(5, "sah", 150)
for it in range(6):
    t0 = time.perf_counter(); d = ctx.upload(s.flat); t1 = time.perf_counter(); d.close(); t2 = time.perf_counter()
    print("upload %.3f ms close %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3), flush=True)
