"""Candidate / violation statistics of the conservative rejection test over whole frames (CPU, oracle-side)."""
import ctypes as C, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle_py as O
O._LIBS["pre"] = (os.path.join(O.HERE, "librt_oracle_pretest.so"), "rt_oracle_run")
lib = C.CDLL(O._LIBS["pre"][0])
def stats():
    v = (C.c_longlong * 7)()
    lib.rt_oracle_pretest_stats(v)
    return list(v)
jobs = [(5, "sah", 150, 400, 300), (5, "rgrid", 150, 400, 300), (5, "kd", 150, 400, 300), (5, "fgrid", 150, 400, 300),
        (4, "sah", 150, 400, 300), (4, "rgrid", 150, 400, 300)]
if len(sys.argv) > 1:
    w, h = int(sys.argv[1]), int(sys.argv[2])
    jobs = [(p, a, s, w, h) for (p, a, s, _, _) in jobs[:int(sys.argv[3]) if len(sys.argv) > 3 else 6]]
for (p, a, s, w, h) in jobs:
    t0 = time.time()
    r = O.run("pre", p, a, s, w, h, image=True)
    st = stats()
    tests, cand, upd, viol, lists, l1, l2 = st
    print(f"preset {p} {a:6s} {w}x{h}: tests {tests} cand {cand} ({100*cand/max(tests,1):.3f}%) updates {upd} VIOLATIONS {viol} "
          f"lists {lists} with>=1 {l1} ({100*l1/max(lists,1):.2f}%) with>=2 {l2} ({100*l2/max(lists,1):.3f}%)  cand/update {cand/max(upd,1):.3f}  [{time.time()-t0:.1f}s]", flush=True)
