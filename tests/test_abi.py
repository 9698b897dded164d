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
    assert lib.rtb_abi_version() == 2
    assert C.sizeof(rtb200.Material) == 64 and C.sizeof(rtb200.Prim) == 48
    assert C.sizeof(rtb200.KdNode) == 8 and C.sizeof(rtb200.CellWord) == 8


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
