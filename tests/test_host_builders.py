"""Host-side logic of the product (no GPU needed): the C++ host API builds every preset scene and
its accelerator and the flattened result is IDENTICAL to what the unmodified reference builds --
same triangle stream (bit for bit), same grid dims / cell lists, same k-d trees and leaf orders --
checked through the canonical structure hashes recorded from the reference in tests/golden/.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import rtb200
from rtb200 import PresetScene

with open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")) as f:
    META = json.load(f)
TUNNEL_JOBS = sorted(k for k, v in META.items() if v["job"]["preset"] >= 4)


@pytest.mark.parametrize("name", TUNNEL_JOBS)
def test_builder_matches_reference(name):
    g = META[name]
    j = g["job"]
    s = PresetScene(j["preset"], j["algorithm"], j["segments"])
    assert s.stats() == g["stats"]
    assert f"{s.tri_hash():016x}" == g["tri_hash"]
    if j["algorithm"] != "linear":
        assert f"{s.struct_hash():016x}" == g["struct_hash"]
    s.close()


def test_exact_grid_binning_matches_the_reference_branch():
    """Tunnel::exactGridBinning bins with Triangle::intersectWithGrid (reference Triangle.cpp:152-199): the grid must be the
    one the reference builds with the branch it keeps compiled out (Tunnel.cpp:435-445) switched on -- fixtures from
    oracle/_ref/libref_sat.so, tests/golden/make_golden.py: sat_golden -- and must hold fewer references than the default."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "sat_golden.json")) as f:
        sat = json.load(f)
    for name, g in sat.items():
        j = g["job"]
        plain = PresetScene(j["preset"], j["algorithm"], j["segments"])
        rtb200.set_exact_grid_binning(True)
        try:
            s = PresetScene(j["preset"], j["algorithm"], j["segments"])
        finally:
            rtb200.set_exact_grid_binning(False)
        assert s.stats() == g["stats"], name
        assert f"{s.struct_hash():016x}" == g["struct_hash"], name
        assert s.stats()["cell_entries"] < plain.stats()["cell_entries"]
        assert s.struct_hash() != plain.struct_hash()
        s.close(); plain.close()


def test_reference_log_structure_kats():
    """Grid Size 400 x 5 x 400 and the leaf counts printed in the reference's logs."""
    s = PresetScene(5, "rgrid", 150)
    st = s.stats()
    assert (st["grid_x"], st["grid_y"], st["grid_z"]) == (400, 5, 400) and st["n_tris"] == 45900
    s.close()
    s = PresetScene(5, "sah", 150)
    st = s.stats()
    assert st["kd_leaves"] == 34478 and st["kd_leaf_refs"] // st["kd_leaves"] == 11
    s.close()
    s = PresetScene(5, "kd", 150)
    st = s.stats()
    assert st["kd_leaves"] == 121941 and st["kd_leaf_refs"] // st["kd_leaves"] == 8
    s.close()


def test_triangles_match_oracle_bitwise():
    from oracle import oracle_py as O
    for preset, seg in [(5, 7), (4, 23)]:
        s = PresetScene(preset, "linear", seg)
        tri, mat = s.triangles()
        o = O.run("oracle", preset, "linear", seg, 8, 6, triangles=True)
        assert np.array_equal(tri.view(np.uint32), o["tri_out"].view(np.uint32))
        assert np.array_equal(mat, o["tri_mat"])
        s.close()


def test_flat_scene_layout_presets_1_to_3():
    s = PresetScene(1)
    f = s.flat.contents
    assert (f.n_prims, f.n_top, f.n_tris, f.n_loose) == (3, 3, 0, 0)
    assert [f.prims[i].type for i in range(3)] == [0, 1, 1]
    assert f.materials[f.prims[0].material].kind == 2  # RadianceChecker ground
    assert s.setting.enable_monte_carlo == 1 and s.setting.termination_depth == 5
    s.close()
    s = PresetScene(3)
    f = s.flat.contents
    # 6 planes + light sphere + ONE run of 528 loose STL triangles (reference GeometrySet.cpp:33-86)
    assert (f.n_prims, f.n_top, f.n_loose) == (8, 535, 528)
    assert (f.prims[7].type, f.prims[7].base_id, f.prims[7].first, f.prims[7].count) == (2, 7, 0, 528)
    glass = f.materials[f.prims[7].material]
    assert glass.refractiveness == 1.0 and abs(glass.refractive_index - 1.46) < 1e-6
    s.close()


def test_sparse_cell_directory_is_consistent():
    s = PresetScene(5, "fgrid", 20)
    f = s.flat.contents
    words = np.ctypeslib.as_array(C.cast(f.grid_words, C.POINTER(C.c_uint32)), shape=(f.n_cellwords, 2))
    pop = np.array([bin(int(b)).count("1") for b in words[:, 0]], np.int64)
    assert np.array_equal(np.concatenate([[0], np.cumsum(pop)[:-1]]), words[:, 1].astype(np.int64))
    assert pop.sum() == f.n_cells_used
    start = np.ctypeslib.as_array(f.grid_cell_start, shape=(f.n_cells_used + 1,))
    assert start[0] == 0 and start[-1] == f.n_cell_refs and np.all(np.diff(start.astype(np.int64)) > 0)
    s.close()


def test_kd_nodes_are_preorder():
    s = PresetScene(4, "sah", 30)
    f = s.flat.contents
    nodes = np.ctypeslib.as_array(C.cast(f.kd_nodes, C.POINTER(C.c_uint32)), shape=(f.n_kd_nodes, 2))

    def walk(i):
        if nodes[i, 1] & 3 == 3:
            return i + 1
        end_left = walk(i + 1)
        assert nodes[i, 1] >> 2 == end_left
        return walk(end_left)

    import sys
    sys.setrecursionlimit(10000)
    assert walk(0) == f.n_kd_nodes
    s.close()


def test_shard_rows_partition_the_frame():
    for h in (300, 2880, 37):
        for world in (1, 2, 4, 8):
            seen = []
            for rank in range(world):
                fr = rtb200.make_frame(64, h, rank=rank, world=world, row_block=8)
                ys = rtb200.shard_row_indices(h, rank, world, 8)
                assert rtb200.shard_rows(fr) == len(ys)
                seen.extend(ys.tolist())
            assert sorted(seen) == list(range(h))


def test_performance_test_program_builders_match_reference():
    """rt::PerformanceTest builds its tunnel with that program's own generator and k-d builder (event-sweep SAH with
    automatic termination): triangle count, grid / tree sizes and structure hashes equal what the compiled
    src/PerformanceTest sources produce (tests/golden/bounce_pt_golden.json, recorded from libref_pt.so)."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "bounce_pt_golden.json")) as f:
        meta = json.load(f)
    cases = {"r2000_s30": (2000.0, 1.5708, 30, 30), "r100_a75_s24x12": (100.0, 1.309, 24, 12), "r5000_s150": (5000.0, 1.5707964, 150, 150)}
    for name, m in sorted(meta.items()):
        case, alg = name.split(".")
        if alg == "fgrid" and case != "r2000_s30":
            continue  # (the 400^3 grid of the small tunnel is 6.9 M occupied cells: slow to hash, nothing new)
        s = rtb200.PerfScene(*cases[case], alg)
        st = s.stats()
        for k in ("n_tris", "grid_x", "grid_y", "grid_z", "cells_nonempty", "cell_entries", "cell_max", "kd_nodes", "kd_leaves", "kd_leaf_refs", "kd_max_depth"):
            assert st[k] == m["stats"][k], (name, k)
        assert f"{s.struct_hash():016x}" == m["struct_hash"], name
        s.close()


def test_convex_tables_of_presets_match_reference():
    """RayTracingOpt's convex tables (400 x 400 cell table, ring normals from the path, order table) built by the host API."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "convex_golden.json")) as f:
        meta = json.load(f)
    for name, m in sorted(meta.items()):
        j = m["job"]
        s = PresetScene(j["preset"], j["algorithm"], j["segments"])
        assert f"{s.struct_hash():016x}" == m["struct_hash"], name
        f_ = s.flat.contents
        assert (f_.cx_table_size, f_.cx_round_bins) == (400, 0)
        s.close()
