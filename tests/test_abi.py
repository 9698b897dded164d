"""The C-ABI library loads without a GPU and exports every symbol include/rtb.h declares; the
product path fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

import rtb200

HEADER = os.path.join(rtb200.ROOT, "include", "rtb.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = C.CDLL(rtb200.CUDA_LIB)
    names = _declared()
    assert set(names) == set(rtb200.ABI_SYMBOLS)
    for n in names:
        assert hasattr(lib, n), n


def test_abi_version_and_struct_sizes():
    lib = rtb200.cuda_lib()
    assert lib.rtb_abi_version() == 5
    assert C.sizeof(rtb200.Frame) == 56  # ABI 4 appended sample_first / sample_count (col_block had taken the tail padding in ABI 3)
    assert C.sizeof(rtb200.Material) == 64 and C.sizeof(rtb200.Prim) == 48
    assert C.sizeof(rtb200.KdNode) == 8 and C.sizeof(rtb200.CellWord) == 8
    # ABI 5 appended arrays_page_locked (it took the tail padding: the size is unchanged)
    assert C.sizeof(rtb200.FlatScene) == 320 and rtb200.FlatScene.arrays_page_locked.offset == 312


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rtb200.RtbError) as e:
        rtb200.Context(0)
    assert "no CPU path" in str(e.value) or "rc=-2" in str(e.value)
    with pytest.raises(rtb200.RtbError):
        rtb200.script_run(1, width=16, height=12, samples=1)


def test_product_never_imports_the_oracle():
    pkg = rtb200.PKG
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle_py" not in text and "rt_oracle" not in text.replace("oracle/rt_oracle.cpp", ""), f


def test_column_block_shards_partition_the_frame():
    """Host arithmetic of the column-block sharding (rtb_frame.col_block): for every row the ranks' column sets are a
    partition of the row, every rank owns width / world columns, and the block -> rank map rotates with the block row."""
    import numpy as np
    for (w, world, rb, cb) in [(256, 4, 8, 16), (3840, 8, 8, 32), (96, 3, 16, 8), (64, 2, 8, 32)]:
        for y in (0, 7, 8, 9, rb * world - 1, rb * world, 1000):
            seen = np.zeros(w, np.int32)
            for rank in range(world):
                xs = rtb200.shard_col_indices(w, y, rank, world, rb, cb)
                assert len(xs) == w // world and np.all(np.diff(xs) > 0)
                assert np.all(((xs // cb) + (y // rb)) % world == rank)
                seen[xs] += 1
            assert np.all(seen == 1)
    lib = rtb200.cuda_lib()
    fr = rtb200.make_frame(3840, 2880, rank=3, world=8, row_block=8, col_block=32)
    assert lib.rtb_shard_rows(C.byref(fr)) == 2880 and lib.rtb_shard_width(C.byref(fr)) == 480
    bad = rtb200.make_frame(3840, 2880, rank=3, world=8, row_block=8, col_block=12)   # not a multiple of 8
    assert lib.rtb_shard_rows(C.byref(bad)) == -1
    bad = rtb200.make_frame(3848, 2880, rank=3, world=8, row_block=8, col_block=32)   # width not a multiple of world * col_block
    assert lib.rtb_shard_rows(C.byref(bad)) == -1
    one = rtb200.make_frame(3840, 2880, rank=0, world=1, row_block=8, col_block=32)   # a single rank ignores col_block
    assert lib.rtb_shard_rows(C.byref(one)) == 2880 and lib.rtb_shard_width(C.byref(one)) == 3840


def test_kd_validate_matches_a_recursive_definition():
    """rtb_kd_validate (the linear pre-order check rtb_scene_upload runs) against the recursive definition of a
    well-formed pre-order k-d array, on random valid trees, targeted corruptions and exhaustive small garbage."""
    import itertools
    import numpy as np
    lib = rtb200.cuda_lib()
    lib.rtb_kd_validate.argtypes = [C.POINTER(rtb200.KdNode), C.c_int32, C.POINTER(C.c_int32)]
    lib.rtb_kd_validate.restype = C.c_int

    def definition(nodes):
        """(valid, max depth): subtree(i) ends where its right subtree ends; the left subtree must end at `right`."""
        n = len(nodes)
        deepest = 0

        def end(i, depth):
            nonlocal deepest
            if i >= n or depth > 64:
                return None
            deepest = max(deepest, depth)
            a, b = nodes[i]
            if b & 3 == 3:
                return i + 1
            right = b >> 2
            left_end = end(i + 1, depth + 1)
            if left_end is None or left_end != right or right >= n:
                return None
            return end(right, depth + 1)

        return end(0, 0) == n, deepest

    def call(nodes):
        arr = (rtb200.KdNode * len(nodes))(*[rtb200.KdNode(a, b) for a, b in nodes])
        d = C.c_int32(-1)
        rc = lib.rtb_kd_validate(arr, len(nodes), C.byref(d))
        return rc == 0, d.value

    rng = np.random.default_rng(3)

    def build(depth_left, base):
        """returns the pre-order node list of a random subtree whose first node has index `base`"""
        if depth_left == 0 or rng.random() < 0.3:
            return [(int(rng.integers(0, 100)), (int(rng.integers(0, 9)) << 2) | 3)]
        left = build(depth_left - 1, base + 1)
        right_index = base + 1 + len(left)
        right = build(depth_left - 1, right_index)
        return [(int(rng.integers(0, 2 ** 31)), (right_index << 2) | int(rng.integers(0, 3)))] + left + right

    for _ in range(200):
        nodes = build(int(rng.integers(1, 12)), 0)
        ok, depth = definition(nodes)
        assert ok and call(nodes) == (True, depth)
        if len(nodes) > 1:
            for _ in range(6):  # one random field damaged
                bad = list(nodes)
                i = int(rng.integers(0, len(bad)))
                a, b = bad[i]
                choice = int(rng.integers(0, 4))
                if choice == 0:
                    b = (int(rng.integers(0, len(bad) + 3)) << 2) | (b & 3)
                elif choice == 1:
                    b = (b & ~3) | int(rng.integers(0, 4))
                elif choice == 2:
                    bad = bad[:-1]
                else:
                    bad = bad + [bad[-1]]
                if choice < 2:
                    bad[i] = (a, b)
                assert call(bad)[0] == definition(bad)[0], (choice, i, bad)
    # exhaustive: every array of up to 5 nodes over {leaf, inner with right in 0..5}
    kinds = [(0, 3)] + [(0, (r << 2) | 0) for r in range(0, 6)]
    for n in range(1, 6):
        for combo in itertools.product(kinds, repeat=n):
            ok, depth = definition(list(combo))
            got = call(list(combo))
            assert got[0] == ok, combo
            if ok:
                assert got[1] == depth
    # a degenerate chain deeper than 64 is refused
    deep = []
    for i in range(70):
        deep.append((0, ((2 * 70 - i) << 2) | 0))      # inner: left child next, right child = a leaf near the end
    deep += [(0, 3)] * 71
    assert not call(deep)[0]
