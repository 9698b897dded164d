import sys
sys.path.insert(0, '.')
import numpy as np, rtb200
from oracle import oracle_py as O
rng = np.random.default_rng(1)
for n in (1000, 100000):
    xy = rng.random((n, 2), dtype=np.float32)
    for alg in ["rgrid", "fgrid", "kd", "sah"]:
        g = rtb200.perf_test(xy, 2000.0, 1.5708, 150, 150, alg)
        g = rtb200.perf_test(xy, 2000.0, 1.5708, 150, 150, alg)
        line = f"N={n} {alg}: GPU trace {g['trace_ms']:.2f} ms, {g['total_rays']} rays, {g['total_rays']/g['trace_ms']/1e3:.1f} Mrays/s, preprocess {g['preprocess_ms']:.0f} ms, all reached {bool(g['reached'].all())}"
        if n == 1000:
            c = O.bounce("ref_timing" if O.available("ref_timing") else "oracle", xy, 2000.0, 1.5708, 150, 150, alg, threads=1)
            line += f" | reference 1 thread: trace {c['trace_ms']:.1f} ms, preprocess {c['prepare_ms']:.0f} ms, same depths {bool(np.array_equal(c['depth'], g['depth']))}"
        print(line, flush=True)
