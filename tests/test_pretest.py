"""The conservative rejection test of the throughput kernels (win32-ray-tracing-demo_b200/csrc/rtb_pretest.h) against the
oracle's exact triangle test, on the CPU.

The header is host + device code: oracle/librt_oracle_pretest.so (make -C oracle pretest) compiles the SAME function
next to the exact Triangle::intersect restatement and evaluates both on every (ray, triangle) pair a frame tests.
A violation is a pair the pre-test rejects although the exact test would have made it the nearest hit of its list --
it must never happen, whatever the scene.  The GPU side of the argument (bit-identical hit ids, distances, counters
and images with the pre-test in the kernels) is tests/test_gpu_parity.py.
"""
import ctypes as C
import os
import subprocess

import pytest

from oracle import oracle_py as O

LIB = os.path.join(O.HERE, "librt_oracle_pretest.so")


@pytest.fixture(scope="module")
def pre():
    subprocess.check_call(["make", "-s", "-C", O.HERE, "pretest"])
    O._LIBS["pretest"] = (LIB, "rt_oracle_run")
    lib = C.CDLL(LIB)
    lib.rt_oracle_pretest_fuzz.argtypes = [C.c_uint64, C.c_longlong, C.POINTER(C.c_longlong)]
    return lib


def _stats(lib):
    v = (C.c_longlong * 8)()
    lib.rt_oracle_pretest_stats(v)
    return dict(zip(["tests", "candidates", "updates", "violations", "lists", "lists_ge1", "lists_ge2", "unsure"], v))


@pytest.mark.parametrize("use_nearest", [0, 1])
@pytest.mark.parametrize("preset,algorithm,segments,width,height", [
    (5, "sah", 150, 160, 120), (5, "kd", 150, 160, 120), (5, "rgrid", 150, 160, 120), (5, "fgrid", 60, 160, 120),
    (4, "sah", 150, 160, 120), (4, "rgrid", 40, 120, 90), (5, "sah", 24, 200, 150), (5, "rgrid", 7, 200, 150),
])
def test_pretest_never_rejects_a_hit(pre, preset, algorithm, segments, width, height, use_nearest):
    """Whole Whitted frames (all bounces): zero violations; the candidates are essentially the real hits."""
    pre.rt_oracle_pretest_mode(use_nearest)
    _stats(pre)  # reset
    r = O.run("pretest", preset, algorithm, segments, width, height, image=True)
    st = _stats(pre)
    pre.rt_oracle_pretest_mode(0)
    assert st["tests"] == r["n_tri_tests"] > 0
    assert st["violations"] == 0
    assert st["updates"] > 0
    # what the deferral in the kernels relies on (speed only, never correctness): a handful of candidates per real
    # hit -- 1.1-1.6 on the 150-segment frames, more where long rays meet the big slivers of a coarse tessellation,
    # whose beta the reference itself computes with a large rounding error -- and a second candidate in a list is rare
    if segments == 150:
        assert st["candidates"] <= 2.5 * st["updates"]
        assert st["lists_ge2"] <= 0.02 * st["lists"]
        assert st["unsure"] <= 0.01 * st["tests"]


def test_pretest_random_stress(pre):
    """Triangles from 1e-3 to 1e4 units, slivers up to 1000:1, scene offsets up to 1e5, rays aimed at the triangle's
    boundary from 0.01 to 1000 triangle sizes away, head-on to grazing, unit and non-unit directions, with and without
    a leaf window around the hit: no accepted hit is ever rejected."""
    out = (C.c_longlong * 4)()
    pre.rt_oracle_pretest_fuzz(20261018, 30_000_000, out)
    pairs, accepts, candidates, violations = list(out)
    assert violations == 0
    assert accepts > 0.15 * pairs          # the stress really exercises accepted hits
    assert candidates >= accepts


@pytest.mark.parametrize("algorithm,pt_builders,radius,angle", [("sah", True, 2000.0, 1.5708), ("rgrid", False, 500.0, 1.0),
                                                                ("kd", False, 2000.0, 1.5708)])
def test_pretest_on_mirror_bounce_chains(pre, algorithm, pt_builders, radius, angle):
    """The PerformanceTest workload (up to 200 mirror bounces per sample down the tunnel: long chains of grazing rays
    started exactly on a surface, the PT program's own ring orientation and event-sweep SAH tree): zero violations."""
    import numpy as np
    rng = np.random.default_rng(5)
    xy = rng.random((3000, 2)).astype(np.float32)
    pre.rt_oracle_pretest_mode(0)
    _stats(pre)  # reset
    O._LIBS["pretest"] = (LIB, "rt_oracle_run")
    r = O.bounce("pretest", xy, radius=radius, angle=angle, arch_seg=150, path_seg=150, algorithm=algorithm, max_depth=200,
                 pt_builders=pt_builders)
    st = _stats(pre)
    assert r["total_rays"] > 3000 and st["tests"] > 0
    assert st["violations"] == 0
    assert st["updates"] > 0


def test_pretest_error_bound_against_exact_determinants(pre):
    """The bound itself: the reference's float determinants, the cheap FMA determinants of rtb_pretest.h and the exact
    ones (binary128) on random pairs.  The header claims |det_ref - Det| <= gamma_7 S, |det' - Det| <= gamma_5 S and
    kappa / kT >= |det' - det_ref| with a factor >= 2.6 to spare; the checker's operation-by-operation restatement of
    the header must decide exactly like the header (no drift between the two)."""
    pre.rt_oracle_pretest_bound_fuzz.argtypes = [C.c_uint64, C.c_longlong, C.POINTER(C.c_double)]
    out = (C.c_double * 5)()
    pre.rt_oracle_pretest_bound_fuzz(7, 4_000_000, out)
    ref_vs_exact, cheap_vs_exact, kappa_ratio, kt_ratio, drift = list(out)
    assert 0 < ref_vs_exact <= 1.0
    assert 0 < cheap_vs_exact <= 1.0
    assert 0 < kappa_ratio <= 1 / 2.6
    assert 0 < kt_ratio <= 1 / 2.6
    assert drift == 0
