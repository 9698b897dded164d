"""Generate tests/golden/ref_golden.{json,npz} from the UNMODIFIED reference (oracle/_ref/libref.so).

Run in the build container, where /root/reference exists:

    make -C oracle ref && python tests/golden/make_golden.py

The reference ships no per-ray or per-pixel golden vectors, so these
fixtures are outputs of the reference itself, compiled headless by oracle/build_ref.sh.  They
travel with the repo to the GPU box, where /root/reference does not exist, and pin both the CPU
restatement (oracle/rt_oracle.cpp) and the CUDA path.

Small jobs keep their full arrays in the .npz; full-size jobs keep sha256 digests + scalars.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_py as O  # noqa: E402

ARRAYS = ["image", "image8", "hit_id", "hit_t", "seq_len", "seq_hash"]

# (name, kwargs, keep_arrays)
JOBS = []
for alg in ["linear", "rgrid", "fgrid", "kd", "sah"]:
    JOBS.append((f"p5_{alg}_s12_80x60", dict(preset=5, algorithm=alg, segments=12, width=80, height=60), True))
    JOBS.append((f"p4_{alg}_s12_80x60", dict(preset=4, algorithm=alg, segments=12, width=80, height=60), True))
for alg in ["rgrid", "fgrid", "kd", "sah"]:
    JOBS.append((f"p5_{alg}_s150_400x300", dict(preset=5, algorithm=alg, segments=150, width=400, height=300), False))
for alg in ["rgrid", "sah"]:
    JOBS.append((f"p4_{alg}_s150_400x300", dict(preset=4, algorithm=alg, segments=150, width=400, height=300), False))
JOBS += [
    ("p5_sah_s40_160x120", dict(preset=5, algorithm="sah", segments=40, width=160, height=120), True),
    ("p5_rgrid_s40_160x120", dict(preset=5, algorithm="rgrid", segments=40, width=160, height=120), True),
    ("p1_mc_80x60_spp4", dict(preset=1, width=80, height=60, samples=4), True),
    ("p2_mc_80x60_spp4", dict(preset=2, width=80, height=60, samples=4), True),
    ("p3_mc_40x30_spp2", dict(preset=3, width=40, height=30, samples=2), True),
    ("p2_hq_40x30_spp3", dict(preset=2, width=40, height=30, samples=3, setting="highquality"), True),
    ("p2_hs_40x30_spp3", dict(preset=2, width=40, height=30, samples=3, setting="highspeed"), True),
    ("p1_simple_80x60", dict(preset=1, width=80, height=60, setting="simple"), True),
    ("p2_simple_80x60", dict(preset=2, width=80, height=60, setting="simple"), True),
    ("p3_simple_40x30", dict(preset=3, width=40, height=30, setting="simple"), True),
]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    meta, arrays = {}, {}
    for name, kw, keep in JOBS:
        r = O.run("ref", image=True, hits=True, seq=True, **kw)
        entry = {"job": kw, "stats": r["stats"], "struct_hash": f"{r['struct_hash']:016x}",
                 "tri_hash": f"{r['tri_hash']:016x}", "n_rays": r["n_rays"], "n_tri_tests": r["n_tri_tests"],
                 "n_steps": r["n_steps"], "sha256": {k: digest(r[k]) for k in ARRAYS}}
        meta[name] = entry
        if keep:
            for k in ARRAYS:
                arrays[f"{name}.{k}"] = r[k]
        print(name, entry["stats"], entry["n_rays"], flush=True)
    with open(os.path.join(HERE, "ref_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **arrays)


if __name__ == "__main__":
    main()


def bounce_golden():
    """tests/golden/bounce_golden.npz: the PerformanceTest ray loop on the reference's classes (ref_bounce)."""
    xy = np.random.default_rng(21).random((300, 2), dtype=np.float32)
    out = {"xy": xy}
    for alg in ("rgrid", "sah"):
        a = O.bounce("ref", xy, 2000.0, 1.5708, 30, 30, alg)
        for k in ("reached", "depth", "last_id", "last_pos"):
            out[f"{alg}.{k}"] = a[k]
    np.savez_compressed(os.path.join(HERE, "bounce_golden.npz"), **out)


PT_CASES = {  # name -> (radius, angle, arch_seg, path_seg): the tunnels of the PerformanceTest program
    "r2000_s30": (2000.0, 1.5708, 30, 30),
    "r100_a75_s24x12": (100.0, 1.309, 24, 12),
}


def bounce_pt_golden():
    """tests/golden/bounce_pt_golden.{npz,json}: the PerformanceTest program itself (libref_pt.so: its generator, its
    GridAcc / KdTreeAcc / ConvexAcc builders incl. the event-sweep SAH, its trace): per-ray results, structure sizes and hashes;
    plus the full-size trees (150 x 150, radius 5000) -- structure only."""
    xy = np.random.default_rng(22).random((300, 2), dtype=np.float32)
    out, meta = {"xy": xy}, {}
    for name, (radius, angle, aseg, pseg) in PT_CASES.items():
        for alg in ("rgrid", "fgrid", "kd", "sah", "convex", "convexsimple"):
            a = O.bounce("ref_pt", xy, radius, angle, aseg, pseg, alg)
            for k in ("reached", "depth", "last_id", "last_pos"):
                out[f"{name}.{alg}.{k}"] = a[k]
            meta[f"{name}.{alg}"] = {"stats": a["stats"], "struct_hash": f"{a['struct_hash']:016x}", "total_rays": a["total_rays"]}
    for alg in ("kd", "sah", "convex", "convexsimple"):
        a = O.bounce("ref_pt", xy[:4], 5000.0, 1.5707964, 150, 150, alg)
        meta[f"r5000_s150.{alg}"] = {"stats": a["stats"], "struct_hash": f"{a['struct_hash']:016x}"}
    np.savez_compressed(os.path.join(HERE, "bounce_pt_golden.npz"), **out)
    with open(os.path.join(HERE, "bounce_pt_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


CONVEX_JOBS = {  # RayTracingOpt's own convex walk (Tunnel::fastIntersect) behind the preset scenes
    "p5_convex_s40_160x120": dict(preset=5, algorithm="convex", segments=40, width=160, height=120),
    "p5_convexsimple_s40_160x120": dict(preset=5, algorithm="convexsimple", segments=40, width=160, height=120),
    "p4_convex_s24_160x120": dict(preset=4, algorithm="convex", segments=24, width=160, height=120),
}


def convex_golden():
    """tests/golden/convex_golden.{npz,json}: presets 4 / 5 rendered by libref.so with Tunnel::Convex / ConvexSimple."""
    arrays, meta = {}, {}
    for name, job in CONVEX_JOBS.items():
        r = O.run("ref", image=True, hits=True, **job)
        for k in ("hit_id", "hit_t", "image"):
            arrays[f"{name}.{k}"] = r[k]
        meta[name] = {"job": job, "n_rays": r["n_rays"], "n_tri_tests": r["n_tri_tests"], "struct_hash": f"{r['struct_hash']:016x}"}
    np.savez_compressed(os.path.join(HERE, "convex_golden.npz"), **arrays)
    with open(os.path.join(HERE, "convex_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


SAT_JOBS = {  # regular grids binned with the reference's exact overlap test (libref_sat.so: Tunnel.cpp:435 enabled)
    "p5_rgrid_s24_96x72": (dict(preset=5, algorithm="rgrid", segments=24, width=96, height=72), True),
    "p4_rgrid_s16_64x48": (dict(preset=4, algorithm="rgrid", segments=16, width=64, height=48), True),
    "p5_rgrid_s150_400x300": (dict(preset=5, algorithm="rgrid", segments=150, width=400, height=300), False),
}


def sat_golden():
    """tests/golden/sat_golden.{npz,json}: the reference rebuilt with its exact grid binning enabled (build_ref.sh P8)."""
    arrays, meta = {}, {}
    for name, (job, keep) in SAT_JOBS.items():
        r = O.run("ref_sat", image=True, hits=True, seq=True, **job)
        meta[name] = {"job": job, "stats": r["stats"], "struct_hash": f"{r['struct_hash']:016x}", "n_rays": r["n_rays"],
                      "n_tri_tests": r["n_tri_tests"], "n_steps": r["n_steps"], "sha256": {k: digest(r[k]) for k in ARRAYS}}
        if keep:
            for k in ARRAYS:
                arrays[f"{name}.{k}"] = r[k]
    np.savez_compressed(os.path.join(HERE, "sat_golden.npz"), **arrays)
    with open(os.path.join(HERE, "sat_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    bounce_golden()
    bounce_pt_golden()
    convex_golden()
    sat_golden()
