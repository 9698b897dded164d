"""The CPU oracle (oracle/rt_oracle.cpp) against the reference's own outputs.

* test_oracle_matches_golden: every committed fixture produced by the unmodified reference
  (tests/golden/make_golden.py) is reproduced bit for bit by the restatement -- structural KATs
  of the reference's logs (grid 400x5x400, 121941 / 34478 leaves), triangle streams, accelerator
  hashes, hit ids, distances, traversal sequences, Whitted and erand48 Monte-Carlo images.
* test_oracle_matches_ref_live: the same comparison against oracle/_ref/libref.so on fresh job
  shapes, when the reference library is present (it is built only where /root/reference exists).
"""
import hashlib
import json

import numpy as np
import pytest

from oracle import oracle_py as O

ARRAYS = ["image", "image8", "hit_id", "hit_t", "seq_len", "seq_hash"]


def _digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _names():
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")) as f:
        return sorted(json.load(f).keys())


@pytest.mark.parametrize("name", _names())
def test_oracle_matches_golden(golden, name):
    meta, arrays = golden
    g = meta[name]
    r = O.run("oracle", image=True, hits=True, seq=True, **g["job"])
    assert r["stats"] == g["stats"]
    assert f"{r['struct_hash']:016x}" == g["struct_hash"]
    assert f"{r['tri_hash']:016x}" == g["tri_hash"]
    assert (r["n_rays"], r["n_tri_tests"], r["n_steps"]) == (g["n_rays"], g["n_tri_tests"], g["n_steps"])
    for k in ARRAYS:
        assert _digest(r[k]) == g["sha256"][k], k
        if f"{name}.{k}" in arrays:
            assert np.array_equal(r[k].view(np.uint8), arrays[f"{name}.{k}"].view(np.uint8)), k


def test_reference_log_kats(golden):
    """Structural values printed in the reference's own logs."""
    meta, _ = golden
    s = meta["p5_rgrid_s150_400x300"]["stats"]      # logs/test-tessellation-150/rgrid.log:7
    assert (s["grid_x"], s["grid_y"], s["grid_z"]) == (400, 5, 400)
    s = meta["p5_fgrid_s150_400x300"]["stats"]      # fgrid.log
    assert (s["grid_x"], s["grid_y"], s["grid_z"]) == (400, 400, 400)
    s = meta["p5_kd_s150_400x300"]["stats"]         # kd-tree.txt:7-8
    assert s["kd_leaves"] == 121941 and s["kd_leaf_refs"] // s["kd_leaves"] == 8
    s = meta["p5_sah_s150_400x300"]["stats"]        # kd-sah.log:7-8
    assert s["kd_leaves"] == 34478 and s["kd_leaf_refs"] // s["kd_leaves"] == 11
    assert meta["p5_sah_s150_400x300"]["stats"]["n_tris"] == 45900


LIVE = [
    dict(preset=5, algorithm="sah", segments=20, width=64, height=48),
    dict(preset=5, algorithm="fgrid", segments=9, width=48, height=36),
    dict(preset=4, algorithm="kd", segments=17, width=64, height=48),
    dict(preset=4, algorithm="rgrid", segments=8, width=48, height=36),
    dict(preset=2, width=32, height=24, samples=5),
    dict(preset=1, width=32, height=24, samples=3, setting="highspeed"),
]


@pytest.mark.skipif(not O.available("ref"), reason="oracle/_ref/libref.so not built (no /root/reference)")
@pytest.mark.parametrize("job", LIVE, ids=lambda j: "-".join(f"{k}{v}" for k, v in j.items()))
def test_oracle_matches_ref_live(job):
    a = O.run("ref", image=True, hits=True, seq=True, seq_cap=16, triangles=job["preset"] >= 4, **job)
    b = O.run("oracle", image=True, hits=True, seq=True, seq_cap=16, triangles=job["preset"] >= 4, **job)
    keys = ARRAYS + ["seq_buf"] + (["tri_out", "tri_mat"] if job["preset"] >= 4 else [])
    for k in keys:
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k
    for k in ["stats", "struct_hash", "tri_hash", "n_rays", "n_tri_tests", "n_steps"]:
        assert a[k] == b[k], k


def test_counter_rng_mode_is_unbiased_vs_erand48():
    """The counter-based stream (what the CUDA path uses) must estimate the same image as the
    reference's erand48 stream: compare image means of preset 1 at 64 spp."""
    a = O.run("oracle", 1, width=48, height=36, samples=64, image=True, rng=0)["image"]
    b = O.run("oracle", 1, width=48, height=36, samples=64, image=True, rng=1, seed=7)["image"]
    assert abs(a.mean() - b.mean()) / a.mean() < 0.02


@pytest.mark.skipif(not O.available("ref"), reason="oracle/_ref/libref.so not built (no /root/reference)")
@pytest.mark.parametrize("alg", ["rgrid", "fgrid", "kd", "sah"])
def test_bounce_workload_oracle_matches_ref(alg):
    """PerformanceTest ray loop (main.cpp:29-59) on the reference's own intersection classes."""
    xy = np.random.default_rng(3).random((150, 2), dtype=np.float32)
    a = O.bounce("ref", xy, 2000.0, 1.5708, 24, 24, alg)
    b = O.bounce("oracle", xy, 2000.0, 1.5708, 24, 24, alg)
    for k in ("reached", "depth", "last_id", "last_pos"):
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k
    assert a["total_rays"] == b["total_rays"] and a["reached"].all()


def test_bounce_workload_golden():
    """Fixture recorded from libref.so (tests/golden/bounce_golden.npz)."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "bounce_golden.npz"))
    for alg in ("rgrid", "sah"):
        b = O.bounce("oracle", g["xy"], 2000.0, 1.5708, 30, 30, alg)
        for k in ("reached", "depth", "last_id", "last_pos"):
            assert np.array_equal(b[k].view(np.uint8), g[f"{alg}.{k}"].view(np.uint8)), (alg, k)


# ---- the PerformanceTest PROGRAM itself (its own tunnel generator and accelerator builders) ----------------
PT_CASES = {"r2000_s30": (2000.0, 1.5708, 30, 30), "r100_a75_s24x12": (100.0, 1.309, 24, 12)}


@pytest.mark.skipif(not O.available("ref_pt"), reason="oracle/_ref/libref_pt.so not built (no /root/reference)")
@pytest.mark.parametrize("alg", ["linear", "rgrid", "fgrid", "kd", "sah", "convex", "convexsimple"])
def test_pt_program_oracle_matches_ref_pt(alg):
    """oracle (pt_builders) == src/PerformanceTest compiled as it is: tunnel generator (TunnelGenerator.cpp:243-279),
    GridAcc, KdTreeAcc median and event-sweep SAH with automatic termination (KdTreeAcc.cpp:38-274), ConvexAcc tables
    and polygon walk with the ray context carried over the bounces (ConvexAcc.cpp) -- structure hash, sizes and every
    per-ray result of main.cpp's trace."""
    xy = np.random.default_rng(5).random((200, 2), dtype=np.float32)
    a = O.bounce("ref_pt", xy, 1000.0, 1.5707964, 40, 40, alg)
    b = O.bounce("oracle", xy, 1000.0, 1.5707964, 40, 40, alg, pt_builders=True)
    assert a["struct_hash"] == b["struct_hash"] and a["stats"] == b["stats"]
    for k in ("reached", "depth", "last_id", "last_pos"):
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k
    assert a["total_rays"] == b["total_rays"] and a["reached"].all()


@pytest.mark.parametrize("case", sorted(PT_CASES))
def test_pt_program_golden(case):
    """Fixtures recorded from libref_pt.so (tests/golden/bounce_pt_golden.{npz,json}, make_golden.py)."""
    import json
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    g = np.load(os.path.join(here, "bounce_pt_golden.npz"))
    with open(os.path.join(here, "bounce_pt_golden.json")) as f:
        meta = json.load(f)
    radius, angle, aseg, pseg = PT_CASES[case]
    for alg in ("rgrid", "kd", "sah", "convex", "convexsimple") + (("fgrid",) if case == "r2000_s30" else ()):
        b = O.bounce("oracle", g["xy"], radius, angle, aseg, pseg, alg, pt_builders=True)
        m = meta[f"{case}.{alg}"]
        assert f"{b['struct_hash']:016x}" == m["struct_hash"] and b["stats"] == m["stats"] and b["total_rays"] == m["total_rays"]
        for k in ("reached", "depth", "last_id", "last_pos"):
            assert np.array_equal(b[k].view(np.uint8), g[f"{case}.{alg}.{k}"].view(np.uint8)), (alg, k)


def test_pt_program_full_size_trees_golden():
    """The 150 x 150 tunnel of radius 5000: median tree 243,773 nodes, event-sweep SAH tree 67,333 nodes / 33,667 leaves /
    372,819 references -- sizes and hashes of the trees src/PerformanceTest builds."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "bounce_pt_golden.json")) as f:
        meta = json.load(f)
    xy = np.full((2, 2), 0.5, np.float32)
    for alg in ("kd", "sah", "convex", "convexsimple"):
        b = O.bounce("oracle", xy, 5000.0, 1.5707964, 150, 150, alg, pt_builders=True)
        m = meta[f"r5000_s150.{alg}"]
        assert f"{b['struct_hash']:016x}" == m["struct_hash"] and b["stats"] == m["stats"]


# ---- RayTracingOpt's convex walk (Tunnel::fastIntersect, Tunnel.cpp:135-344, 972-1161) behind the presets -----
def _convex_meta():
    import json
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    with open(os.path.join(here, "convex_golden.json")) as f:
        return json.load(f), np.load(os.path.join(here, "convex_golden.npz"))


@pytest.mark.parametrize("name", ["p5_convex_s40_160x120", "p5_convexsimple_s40_160x120", "p4_convex_s24_160x120"])
def test_convex_presets_golden(name):
    """Fixtures recorded from libref.so: table hash, primary hit ids / distances, Whitted image (the ray context travels
    down the reflection chain), ray and triangle-test totals -- all bit-identical."""
    meta, arr = _convex_meta()
    m = meta[name]
    o = O.run("oracle", image=True, hits=True, **m["job"])
    assert f"{o['struct_hash']:016x}" == m["struct_hash"]
    assert (o["n_rays"], o["n_tri_tests"]) == (m["n_rays"], m["n_tri_tests"])
    assert np.array_equal(o["hit_id"], arr[f"{name}.hit_id"])
    assert np.array_equal(o["hit_t"].view(np.uint32), arr[f"{name}.hit_t"].view(np.uint32))
    assert np.array_equal(o["image"].view(np.uint32), arr[f"{name}.image"].view(np.uint32))


@pytest.mark.skipif(not O.available("ref"), reason="oracle/_ref/libref.so not built (no /root/reference)")
@pytest.mark.parametrize("alg", ["convex", "convexsimple"])
def test_convex_presets_oracle_matches_ref(alg):
    job = dict(preset=5, algorithm=alg, segments=30, width=96, height=72)
    a = O.run("ref", image=True, hits=True, **job)
    b = O.run("oracle", image=True, hits=True, **job)
    assert a["struct_hash"] == b["struct_hash"] and a["n_rays"] == b["n_rays"]
    for k in ("hit_id", "hit_t", "image"):
        assert np.array_equal(a[k].view(np.uint8), b[k].view(np.uint8)), k


# ---- exact grid binning: the branch the reference carries compiled out (Tunnel.cpp:435-445) -------------------------
def _sat_golden():
    import os
    here = os.path.join(os.path.dirname(__file__), "golden")
    with open(os.path.join(here, "sat_golden.json")) as f:
        return json.load(f), np.load(os.path.join(here, "sat_golden.npz"))


@pytest.mark.parametrize("name", ["p5_rgrid_s24_96x72", "p4_rgrid_s16_64x48", "p5_rgrid_s150_400x300"])
def test_oracle_exact_grid_binning_matches_the_reference_branch(name):
    """oracle_job.grid_exact restates Triangle::intersectWithGrid (Triangle.cpp:152-199).  Fixtures: the reference rebuilt
    with the guarded branch enabled (oracle/_ref/libref_sat.so, build_ref.sh patch P8): grid statistics, structure hash,
    per-ray hit ids / distances / cell sequences, image and counts must all be reproduced bit for bit."""
    meta, arrays = _sat_golden()
    g = meta[name]
    r = O.run("oracle", image=True, hits=True, seq=True, grid_exact=True, **g["job"])
    assert r["stats"] == g["stats"]
    assert f"{r['struct_hash']:016x}" == g["struct_hash"]
    assert (r["n_rays"], r["n_tri_tests"], r["n_steps"]) == (g["n_rays"], g["n_tri_tests"], g["n_steps"])
    for k in ARRAYS:
        assert _digest(r[k]) == g["sha256"][k], k
        if f"{name}.{k}" in arrays.files:
            assert np.array_equal(np.ascontiguousarray(r[k]).view(np.uint8), np.ascontiguousarray(arrays[f"{name}.{k}"]).view(np.uint8)), k
    plain = O.run("oracle", **g["job"])
    assert plain["stats"]["cell_entries"] > r["stats"]["cell_entries"]  # fewer references per cell: what the option is for


def test_oracle_exact_grid_binning_live():
    if not O.available("ref_sat"):
        pytest.skip("oracle/_ref/libref_sat.so is built only where the reference sources exist")
    job = dict(preset=5, algorithm="rgrid", segments=31, width=72, height=54)
    a = O.run("ref_sat", image=True, hits=True, seq=True, **job)
    b = O.run("oracle", image=True, hits=True, seq=True, grid_exact=True, **job)
    assert a["stats"] == b["stats"] and a["struct_hash"] == b["struct_hash"]
    for k in ARRAYS:
        assert np.array_equal(np.ascontiguousarray(a[k]).view(np.uint8), np.ascontiguousarray(b[k]).view(np.uint8)), k
