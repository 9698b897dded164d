"""Route A of INTEGRATION.md, compiled and run: the REFERENCE's own classes (its GeometrySet, Tunnel with the grid / k-d
tree ITS builders produce, its materials and its Scripts.cpp, compiled from /root/reference by oracle/build_ref.sh)
+ oracle/ref/route_a.cpp (SceneFlattener + the CudaRender RenderProc) + librtb200.so = oracle/_ref/libroute_a.so.
`scripts[i]->Run(CudaRender, algorithm, AddLog, UpdateProgress, ...)` -- the call of MainWindow.cpp:369 with the
RenderProc swapped -- must give the image the reference's own Render gives (libref.so, same process), within 1e-5
relative, with the same ray / test / step counts."""
import ctypes as C
import os

import numpy as np
import pytest

import rtb200
from oracle import oracle_py as O  # checker only

pytestmark = pytest.mark.gpu

LIB = os.path.join(rtb200.ROOT, "oracle", "_ref", "libroute_a.so")
ALG = rtb200.ALGORITHMS


@pytest.fixture(scope="module")
def route_a():
    if not os.path.exists(LIB):
        pytest.skip("oracle/_ref/libroute_a.so was not built (needs the reference sources at build time)")
    rtb200.cuda_lib()  # librtb200.so first, with global symbols: libroute_a.so links against it
    lib = C.CDLL(LIB)
    lib.route_a_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_void_p,
                                C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong), C.POINTER(C.c_int), C.c_char_p, C.c_int]
    yield lib
    lib.route_a_shutdown()


def run(lib, preset, algorithm="linear", segments=24, w=160, h=120, spp=1, n_devices=1):
    rgb = np.zeros((w, h, 3), np.float32)
    prep, exe = C.c_int(0), C.c_int(0)
    counts, prog = (C.c_longlong * 3)(), (C.c_int * 2)()
    log = C.create_string_buffer(4096)
    rc = lib.route_a_run(preset, ALG[algorithm], segments, w, h, spp, n_devices, rtb200.stl_fixture().encode(), rgb.ctypes.data,
                         C.byref(prep), C.byref(exe), counts, prog, log, 4096)
    return dict(rc=rc, image=np.ascontiguousarray(rgb.transpose(1, 0, 2)), prepare_ms=prep.value, exec_ms=exe.value,
                n_rays=counts[0], n_tri_tests=counts[1], n_steps=counts[2], progress_calls=prog[0], progress_last=prog[1],
                log=log.value.decode(errors="replace"))


@pytest.mark.parametrize("preset,alg", [(5, "sah"), (5, "kd"), (5, "rgrid"), (5, "fgrid"), (5, "linear"), (4, "sah"), (4, "rgrid")])
def test_reference_scripts_render_through_the_gpu(route_a, preset, alg):
    seg, w, h = (8, 80, 60) if alg == "linear" else (24, 160, 120)
    r = run(route_a, preset, alg, seg, w, h)
    assert r["rc"] == 0, r["log"]
    ref = O.run("ref", preset, alg, seg, w, h, image=True)
    err = np.abs(r["image"] - ref["image"])
    assert (err <= 1e-5 * np.abs(ref["image"]) + 1e-7).all(), err.max()
    assert r["n_rays"] == ref["n_rays"]
    assert r["exec_ms"] >= 1 and r["progress_last"] == h and r["progress_calls"] >= 1


def test_reference_monte_carlo_scripts(route_a):
    """Presets 1-3 (Default(): Monte Carlo) through the same binding: different random streams, so the image mean only
    (the statistical criteria live in test_gpu_parity.py); preset 3 reads the STL mesh with the reference's own loader."""
    for preset, w, h, spp, tol in ((1, 160, 120, 16, 0.03), (2, 160, 120, 16, 0.06), (3, 64, 48, 16, 0.15)):
        r = run(route_a, preset, "linear", 0, w, h, spp)
        assert r["rc"] == 0, r["log"]
        ref = O.run("ref", preset, width=w, height=h, samples=spp, image=True)
        assert np.isfinite(r["image"]).all()
        # a few thousand pixels at 16 spp: the image mean of smallpt carries a few per cent of noise (two CPU seeds differ as much)
        rel = abs(r["image"].mean() - ref["image"].mean()) / ref["image"].mean()
        assert rel <= tol, (preset, rel)


def test_route_a_on_all_devices(route_a):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    one = run(route_a, 5, "sah", 40, 640, 480)
    many = run(route_a, 5, "sah", 40, 640, 480, n_devices=n)
    assert many["rc"] == 0, many["log"]
    assert np.array_equal(one["image"].view(np.uint32), many["image"].view(np.uint32)) and one["n_rays"] == many["n_rays"]


def test_failure_goes_to_the_log_callback(route_a):
    r = run(route_a, 5, "sah", 8, 0, 10)
    assert r["rc"] < 0 and "render" in r["log"].lower() or "image size" in r["log"]
