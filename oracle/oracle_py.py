"""ctypes front-end for the two CPU checkers (oracle/oracle_abi.h).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

  run("oracle", ...)      -> oracle/librt_oracle.so   (CPU restatement, rt_oracle.cpp)
  run("ref", ...)         -> oracle/_ref/libref.so    (the unmodified reference, hooks on)
  run("ref_timing", ...)  -> oracle/_ref/libref_timing.so (hooks compiled out; for timing)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STAT_NAMES = ["n_top", "n_tris", "grid_x", "grid_y", "grid_z", "cells_nonempty", "cell_entries",
              "cell_max", "kd_nodes", "kd_leaves", "kd_leaf_refs", "kd_max_depth"]
ALGORITHMS = {"linear": 0, "rgrid": 1, "fgrid": 2, "kd": 3, "sah": 4, "convex": 5, "convexsimple": 6}
SETTINGS = {"preset": 0, "simple": 1, "default": 2, "highspeed": 3, "highquality": 4}


class OracleJob(C.Structure):
    _fields_ = [
        ("preset", C.c_int32), ("algorithm", C.c_int32), ("segments", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32), ("samples", C.c_int32),
        ("setting", C.c_int32), ("threads", C.c_int32), ("rng", C.c_int32),
        ("seq_cap", C.c_int32), ("tri_cap", C.c_int32), ("repeat", C.c_int32),
        ("seed", C.c_uint64), ("stl_path", C.c_char_p),
        ("rgb", C.c_void_p), ("hit_id", C.c_void_p), ("hit_t", C.c_void_p),
        ("seq_len", C.c_void_p), ("seq_hash", C.c_void_p), ("seq_buf", C.c_void_p),
        ("tri_out", C.c_void_p), ("tri_mat", C.c_void_p),
        ("n_rays", C.c_int64), ("n_tri_tests", C.c_int64), ("n_steps", C.c_int64),
        ("render_ms", C.c_double), ("prepare_ms", C.c_double),
        ("stats", C.c_int64 * 16), ("struct_hash", C.c_uint64), ("tri_hash", C.c_uint64),
        ("render_ms_all", C.c_void_p), ("rgb8", C.c_void_p), ("moments", C.c_void_p),
        ("grid_exact", C.c_int32), ("pad2_", C.c_int32),
    ]


_LIBS = {
    "oracle": (os.path.join(HERE, "librt_oracle.so"), "rt_oracle_run"),
    "ref": (os.path.join(HERE, "_ref", "libref.so"), "ref_run"),
    "ref_timing": (os.path.join(HERE, "_ref", "libref_timing.so"), "ref_run"),
    "ref_pt": (os.path.join(HERE, "_ref", "libref_pt.so"), "ref_pt_bounce"),
    # the reference with the exact grid binning it carries compiled out (Tunnel.cpp:435-445) switched on: build_ref.sh patch P8
    "ref_sat": (os.path.join(HERE, "_ref", "libref_sat.so"), "ref_run"),
}
_loaded = {}


def stl_fixture():
    return os.path.join(os.path.dirname(HERE), "win32-ray-tracing-demo_b200", "assets", "ball.stl")  # input data of preset 3 (the package's asset)


def available(which):
    return os.path.exists(_LIBS[which][0])


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", HERE, "librt_oracle.so"])


def _fn(which):
    if which not in _loaded:
        path, sym = _LIBS[which]
        if which == "oracle" and not os.path.exists(path):
            build_oracle()
        lib = C.CDLL(path)
        fn = getattr(lib, sym)
        fn.argtypes = [C.POINTER(OracleJob)]
        fn.restype = C.c_int
        _loaded[which] = fn
    return _loaded[which]


def run(which, preset, algorithm="linear", segments=150, width=400, height=300, samples=1,
        setting="preset", threads=0, rng=0, seed=0, image=False, hits=False, seq=False, seq_cap=0,
        triangles=False, repeat=0, stl_path=None, moments=False, grid_exact=False):
    """Run one job; returns a dict of numpy arrays / scalars."""
    job = OracleJob()
    job.preset = preset
    job.algorithm = ALGORITHMS[algorithm] if isinstance(algorithm, str) else int(algorithm)
    job.segments, job.width, job.height, job.samples = segments, width, height, samples
    job.setting = SETTINGS[setting] if isinstance(setting, str) else int(setting)
    job.threads, job.rng, job.seed, job.repeat = threads, rng, seed, repeat
    job.stl_path = (stl_path or stl_fixture()).encode()
    job.grid_exact = 1 if grid_exact else 0  # librt_oracle.so; libref_sat.so bins exactly whatever this says
    n = width * height
    keep = {}

    def out(name, arr):
        keep[name] = arr
        setattr(job, name, arr.ctypes.data)

    if image:
        out("rgb", np.zeros((width, height, 3), np.float32))  # reference order: [x][y]
        out("rgb8", np.zeros((width, height, 3), np.uint8))
    if moments:  # librt_oracle.so only: per-pixel sum and sum of squares of the per-sample radiance
        out("moments", np.zeros((width, height, 6), np.float64))
    if hits:
        out("hit_id", np.full(n, -2, np.int32))
        out("hit_t", np.zeros(n, np.float32))
    if seq:
        out("seq_len", np.zeros(n, np.int32))
        out("seq_hash", np.zeros(n, np.uint64))
        if seq_cap:
            job.seq_cap = seq_cap
            out("seq_buf", np.zeros((n, seq_cap), np.int32))
    if triangles:
        cap = 2 * (segments + 3) * segments + 16
        job.tri_cap = cap
        out("tri_out", np.zeros((cap, 12), np.float32))
        out("tri_mat", np.zeros(cap, np.int32))
    if repeat > 0:
        out("render_ms_all", np.zeros(repeat, np.float64))
    rc = _fn(which)(C.byref(job))
    if rc != 0:
        raise RuntimeError(f"{which} job failed rc={rc}")
    res = dict(keep)
    res["stats"] = {k: int(job.stats[i]) for i, k in enumerate(STAT_NAMES)}
    if triangles:
        nt = res["stats"]["n_tris"]
        res["tri_out"] = res["tri_out"][:nt]
        res["tri_mat"] = res["tri_mat"][:nt]
    if image:
        res["image"] = np.ascontiguousarray(res["rgb"].transpose(1, 0, 2))  # [y][x][3]
        res["image8"] = np.ascontiguousarray(res["rgb8"].transpose(1, 0, 2))
    if moments:
        res["moments"] = np.ascontiguousarray(res["moments"].transpose(1, 0, 2))  # [y][x][6]
    for k in ("n_rays", "n_tri_tests", "n_steps", "render_ms", "prepare_ms", "struct_hash", "tri_hash"):
        res[k] = getattr(job, k)
    return res


class BounceJob(C.Structure):
    _fields_ = [("radius", C.c_float), ("angle", C.c_float), ("arch_seg", C.c_int32), ("path_seg", C.c_int32),
                ("algorithm", C.c_int32), ("n", C.c_int32), ("max_depth", C.c_int32), ("threads", C.c_int32),
                ("xy", C.c_void_p), ("reached", C.c_void_p), ("depth", C.c_void_p), ("last_id", C.c_void_p),
                ("last_pos", C.c_void_p), ("total_rays", C.c_int64), ("trace_ms", C.c_double), ("prepare_ms", C.c_double),
                ("pt_builders", C.c_int32), ("pad_", C.c_int32), ("stats", C.c_int64 * 16), ("struct_hash", C.c_uint64)]


def bounce(which, xy, radius=2000.0, angle=1.5708, arch_seg=150, path_seg=150, algorithm="sah", max_depth=200, threads=0,
           pt_builders=False):
    """PerformanceTest workload (oracle_abi.h: oracle_bounce_job) through 'oracle', 'ref' / 'ref_timing' (RayTracingOpt
    classes) or 'ref_pt' (the PerformanceTest sources themselves; always its own builders)."""
    path, _ = _LIBS[which]
    if which == "oracle" and not os.path.exists(path):
        build_oracle()
    lib = C.CDLL(path)
    fn = getattr(lib, {"oracle": "rt_oracle_bounce", "pretest": "rt_oracle_bounce", "ref_pt": "ref_pt_bounce"}.get(which, "ref_bounce"))
    fn.argtypes = [C.POINTER(BounceJob)]
    fn.restype = C.c_int
    xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    n = xy.shape[0]
    out = {"reached": np.zeros(n, np.int32), "depth": np.zeros(n, np.int32), "last_id": np.zeros(n, np.int32),
           "last_pos": np.zeros((n, 3), np.float32)}
    job = BounceJob(radius, angle, arch_seg, path_seg, ALGORITHMS[algorithm] if isinstance(algorithm, str) else algorithm,
                    n, max_depth, threads, xy.ctypes.data, out["reached"].ctypes.data, out["depth"].ctypes.data,
                    out["last_id"].ctypes.data, out["last_pos"].ctypes.data, 0, 0.0, 0.0, 1 if pt_builders else 0, 0)
    rc = fn(C.byref(job))
    if rc != 0:
        raise RuntimeError(f"{which} bounce job failed rc={rc}")
    out.update(total_rays=job.total_rays, trace_ms=job.trace_ms, prepare_ms=job.prepare_ms, struct_hash=int(job.struct_hash),
               stats={k: int(job.stats[i]) for i, k in enumerate(STAT_NAMES)})
    return out
