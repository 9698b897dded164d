import sys
sys.path.insert(0, '.')
import numpy as np, rtb200
from oracle import oracle_py as O
rng = np.random.default_rng(1)
xy = rng.random((100000, 2), dtype=np.float32)
for alg in ["sah", "convexsimple", "convex"]:
    g = rtb200.perf_test(xy, 2000.0, 1.5708, 150, 150, alg)
    g = rtb200.perf_test(xy, 2000.0, 1.5708, 150, 150, alg)
    print(f"N=100000 {alg}: GPU trace {g['trace_ms']:.2f} ms, {g['total_rays']} rays, {g['total_rays']/g['trace_ms']/1e3:.1f} Mrays/s, preprocess {g['preprocess_ms']:.0f} ms, reached {g['reached'].mean():.4f}", flush=True)
    c = O.bounce("ref_pt", xy[:2000], 2000.0, 1.5708, 150, 150, alg, threads=1)
    gg = rtb200.perf_test(xy[:2000], 2000.0, 1.5708, 150, 150, alg)
    print(f"   reference program, 1 thread, 2000 rays: trace {c['trace_ms']:.1f} ms ({c['total_rays']/c['trace_ms']/1e3:.2f} Mrays/s), preprocess {c['prepare_ms']:.0f} ms; GPU identical: {all(np.array_equal(c[k], gg[k]) for k in ('reached','depth','last_id'))}", flush=True)
