"""GPU parity tests (run on the B200 box): every call goes through the C ABI of include/rtb.h
(librtb200.so) or through the C++ host API on top of it, and is compared with

  * the committed outputs of the UNMODIFIED reference (tests/golden/, generated from
    oracle/_ref/libref.so), and
  * the CPU oracle (oracle/rt_oracle.cpp) run live on fresh sizes.

Bars: hit ids, hit distances and traversal sequences per primary ray are BIT-EXACT; deterministic
Whitted images agree within 1e-5 relative (in fact bit-exact wherever no libm powf is involved);
Monte-Carlo images agree path-for-path with the oracle's counter-RNG mode on almost every pixel and
statistically with the reference's erand48 images.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

import rtb200
from rtb200 import PresetScene

pytestmark = pytest.mark.gpu

from oracle import oracle_py as O  # noqa: E402  (checker only)

with open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.json")) as f:
    META = json.load(f)
ARR = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))
KEPT = sorted({k.split(".")[0] for k in ARR.files})
WHITTED_KEPT = [k for k in KEPT if "_mc_" not in k and "_hq_" not in k and "_hs_" not in k]
RTOL = 1e-5  # north_star: deterministic Whitted pixels within 1e-5 relative in float32


@pytest.fixture(scope="module")
def ctx():
    c = rtb200.Context(0)
    yield c
    c.close()


def _scene(ctx, job):
    s = PresetScene(job["preset"], job.get("algorithm", "linear"), job.get("segments", 150))
    return s, ctx.upload(s.flat)


def _setting(s, job):
    return rtb200.make_setting(job["setting"]) if "setting" in job else s.setting


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _assert_image_close(img, ref, what):
    assert img.shape == ref.shape, what
    assert np.all(np.isfinite(img)), what
    err = np.abs(img - ref)
    tol = RTOL * np.abs(ref) + 1e-7
    bad = err > tol
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} components off by up to {err.max()}"


@pytest.mark.parametrize("name", KEPT)
def test_primary_rays_bit_exact_vs_reference(ctx, name):
    """hit id, hit distance and the traversal cell / node sequence of every primary ray."""
    job = META[name]["job"]
    s, dev = _scene(ctx, job)
    r = dev.trace_primary(s.camera, job["width"], job["height"], seq=True)
    assert np.array_equal(r["hit_id"], ARR[f"{name}.hit_id"])
    assert np.array_equal(_bits(r["hit_t"]), _bits(ARR[f"{name}.hit_t"]))
    assert np.array_equal(r["seq_len"], ARR[f"{name}.seq_len"])
    assert np.array_equal(r["seq_hash"], ARR[f"{name}.seq_hash"])
    dev.close(); s.close()


@pytest.mark.parametrize("name", WHITTED_KEPT)
def test_whitted_image_vs_reference(ctx, name):
    job = META[name]["job"]
    s, dev = _scene(ctx, job)
    fr = rtb200.make_frame(job["width"], job["height"], counters=1)
    img, st = dev.render(s.camera, _setting(s, job), fr)
    ref = ARR[f"{name}.image"]
    _assert_image_close(img, ref, name)
    if job["preset"] >= 4 or job["preset"] == 1:
        # chain scenes are folded innermost-first like the recursion: bit-exact except where the
        # Phong ball's powf (glibc vs CUDA libm, a few ulp) is visible
        # (it fills more of the frame in preset 4, where it sits 100 units from the camera)
        same = (_bits(img).reshape(-1, 4) == _bits(ref).reshape(-1, 4)).all(axis=1).mean()
        assert same > (0.97 if job["preset"] == 4 else 0.995), same
    assert st["n_rays"] == META[name]["n_rays"]
    if job["preset"] >= 4:
        assert st["n_tri_tests"] == META[name]["n_tri_tests"]
        assert st["n_steps"] == META[name]["n_steps"]
    dev.close(); s.close()


FULL = ["p5_sah_s150_400x300", "p5_rgrid_s150_400x300", "p5_kd_s150_400x300", "p5_fgrid_s150_400x300",
        "p4_sah_s150_400x300", "p4_rgrid_s150_400x300"]


def _checker(job, **outputs):
    """The CPU checker for a job: the compiled reference itself when oracle/_ref travelled with the repo, else the
    restatement (tests/test_oracle.py pins the two to each other bit for bit)."""
    return O.run("ref" if O.available("ref") else "oracle", **outputs, **job)


@pytest.mark.parametrize("name", FULL)
def test_full_size_preset_digests(ctx, name):
    """BASELINE configs at the demo's default resolution (400x300, 150 segments): sha256 of the per-ray arrays must equal
    the reference's; ray / triangle-test / step counts of the whole frame must be identical; and the Whitted IMAGE is
    compared pixel by pixel at 1e-5 relative with the CPU checker's, whose own image must hash to the committed digest of
    the reference's image (so the comparison is against the reference's pixels, without 8 MB of floats in the repo)."""
    import hashlib
    g = META[name]
    job = g["job"]
    s, dev = _scene(ctx, job)
    r = dev.trace_primary(s.camera, job["width"], job["height"], seq=True)
    for k in ("hit_id", "hit_t", "seq_len", "seq_hash"):
        assert hashlib.sha256(np.ascontiguousarray(r[k]).tobytes()).hexdigest() == g["sha256"][k], k
    fr = rtb200.make_frame(job["width"], job["height"], counters=1)
    img, st = dev.render(s.camera, s.setting, fr)
    assert (st["n_rays"], st["n_tri_tests"], st["n_steps"]) == (g["n_rays"], g["n_tri_tests"], g["n_steps"])
    ref = _checker(job, image=True)
    assert hashlib.sha256(np.ascontiguousarray(ref["image"]).tobytes()).hexdigest() == g["sha256"]["image"]
    _assert_image_close(img, ref["image"], name)
    same = (_bits(img).reshape(-1, 4) == _bits(ref["image"]).reshape(-1, 4)).all(axis=1).mean()
    assert same > (0.97 if job["preset"] == 4 else 0.995), same  # bit-exact except under the Phong ball's powf
    # the second and third frame of the view come from the heaviest-first order and the latency tiers
    for _ in range(2):
        again, st2 = dev.render(s.camera, s.setting, fr)
        assert np.array_equal(_bits(again), _bits(img)) and st2["n_rays"] == st["n_rays"]
    # ... and the 8-bit output stage (MainWindow.cpp:305-311) against the reference's bytes
    img8, _ = dev.render(s.camera, s.setting, rtb200.make_frame(job["width"], job["height"], layout=rtb200.OUTPUT_RGB8))
    assert (img8 != ref["image8"]).mean() < 1e-4  # a byte can differ where a float within 1e-5 sits on a quantisation step
    dev.close(); s.close()


LIVE = [dict(preset=5, algorithm="sah", segments=25, width=96, height=72),
        dict(preset=5, algorithm="kd", segments=33, width=72, height=54),
        dict(preset=4, algorithm="fgrid", segments=10, width=64, height=48),
        dict(preset=4, algorithm="rgrid", segments=14, width=64, height=48),
        dict(preset=5, algorithm="linear", segments=6, width=40, height=30),
        dict(preset=1, width=52, height=39, setting="simple"),
        dict(preset=2, width=52, height=39, setting="simple")]


@pytest.mark.parametrize("job", LIVE, ids=lambda j: "-".join(f"{k}{v}" for k, v in j.items()))
def test_vs_oracle_live(ctx, job):
    """Fresh shapes against the CPU oracle run in the same process."""
    o = O.run("oracle", image=True, hits=True, seq=True, seq_cap=12, **job)
    s, dev = _scene(ctx, job)
    r = dev.trace_primary(s.camera, job["width"], job["height"], seq=True, seq_cap=12)
    assert np.array_equal(r["hit_id"], o["hit_id"])
    assert np.array_equal(_bits(r["hit_t"]), _bits(o["hit_t"]))
    assert np.array_equal(r["seq_len"], o["seq_len"])
    assert np.array_equal(r["seq_hash"], o["seq_hash"])
    assert np.array_equal(r["seq_buf"], o["seq_buf"])
    img, st = dev.render(s.camera, _setting(s, job), rtb200.make_frame(job["width"], job["height"]))
    _assert_image_close(img, o["image"], str(job))
    assert st["n_rays"] == o["n_rays"]
    dev.close(); s.close()


def test_reference_layout_is_column_major(ctx):
    s, dev = _scene(ctx, dict(preset=5, algorithm="sah", segments=12))
    a, _ = dev.render(s.camera, s.setting, rtb200.make_frame(80, 60))
    b, _ = dev.render(s.camera, s.setting, rtb200.make_frame(80, 60, layout=rtb200.LAYOUT_REFERENCE))
    assert b.shape == (80, 60, 3)  # index = x*height + y, reference MainWindow.cpp:276
    assert np.array_equal(a, b.transpose(1, 0, 2))
    dev.close(); s.close()


@pytest.mark.parametrize("world", [2, 3, 8])
def test_row_shards_reassemble_bit_exact(ctx, world):
    """Tile-row sharding (what each GPU of an N-GPU run renders) is bit-identical to the whole frame."""
    w, h = 160, 120
    s, dev = _scene(ctx, dict(preset=5, algorithm="rgrid", segments=40))
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    out = np.zeros_like(whole)
    rays = 0
    for rank in range(world):
        fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8)
        part, pst = dev.render(s.camera, s.setting, fr)
        out[rtb200.shard_row_indices(h, rank, world, 8)] = part
        rays += pst["n_rays"]
    assert np.array_equal(_bits(out), _bits(whole))
    assert rays == st["n_rays"]
    dev.close(); s.close()


@pytest.mark.parametrize("world,col_block,w,h,alg", [(2, 8, 160, 120, "rgrid"), (4, 16, 256, 96, "sah"), (8, 8, 192, 144, "sah"),
                                                     (3, 16, 96, 72, "kd"), (8, 32, 512, 60, "fgrid")])
def test_column_block_shards_reassemble_bit_exact(ctx, world, col_block, w, h, alg):
    """Column-block sharding (rtb_frame.col_block: every rank renders every row and 1 / world of the column blocks,
    rotated from block row to block row) is bit-identical to the whole frame -- second frames too, when the tile order
    and the latency tiers are in play -- and the ray counts add up."""
    s, dev = _scene(ctx, dict(preset=5, algorithm=alg, segments=24))
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    for repeat in range(2):
        out = np.zeros_like(whole)
        rays = 0
        for rank in range(world):
            fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8, col_block=col_block)
            assert rtb200.shard_rows(fr) == h and rtb200.shard_width(fr) == w // world
            part, pst = dev.render(s.camera, s.setting, fr)
            part, pst = dev.render(s.camera, s.setting, fr)  # the second frame of a view uses the recorded tile order
            assert part.shape == (h, w // world, 3)
            for y in range(h):
                out[y, rtb200.shard_col_indices(w, y, rank, world, 8, col_block)] = part[y]
            rays += pst["n_rays"]
        assert np.array_equal(_bits(out), _bits(whole))
        assert rays == st["n_rays"]
    # a width that is not a multiple of world * col_block is refused
    with pytest.raises(rtb200.RtbError):
        dev.render(s.camera, s.setting, rtb200.make_frame(w + 8, h, rank=0, world=world, row_block=8, col_block=col_block))
    dev.close(); s.close()


def test_column_block_shards_monte_carlo_and_rgb8(ctx):
    """The counter RNG is keyed by the GLOBAL pixel, so a column-sharded Monte-Carlo frame equals the whole frame; the
    8-bit output stage goes through the same local layout."""
    w, h, world, cb = 128, 48, 4, 8
    s, dev = _scene(ctx, dict(preset=2))
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=3, seed=11))
    whole8, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=3, seed=11, layout=rtb200.OUTPUT_RGB8))
    out, out8 = np.zeros_like(whole), np.zeros_like(whole8)
    for rank in range(world):
        part, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=3, seed=11, rank=rank, world=world, col_block=cb))
        part8, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=3, seed=11, rank=rank, world=world, col_block=cb,
                                                                      layout=rtb200.OUTPUT_RGB8))
        for y in range(h):
            xs = rtb200.shard_col_indices(w, y, rank, world, 8, cb)
            out[y, xs] = part[y]
            out8[y, xs] = part8[y]
    assert np.array_equal(_bits(out), _bits(whole))
    assert np.array_equal(out8, whole8)
    dev.close(); s.close()


def test_unshard_cols_kernel(ctx):
    import torch
    w, h, world, rb, cb = 256, 72, 4, 8, 16
    s, dev = _scene(ctx, dict(preset=4, algorithm="sah", segments=12))
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    gathered = torch.zeros((world, h, w // world, 3), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()  # the fill above ran on torch's default stream; the handle 0 below selects the library's own stream
    stream = torch.cuda.current_stream().cuda_stream
    for rank in range(world):
        fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=rb, col_block=cb)
        dev.render_device(s.camera, s.setting, fr, gathered[rank].data_ptr(), stream)
    image = torch.empty((h, w, 3), dtype=torch.float32, device="cuda:0")
    rtb200.unshard_cols_device(ctx, gathered.data_ptr(), image.data_ptr(), w, h, world, rb, cb, stream)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(image.cpu().numpy()), _bits(whole))
    dev.close(); s.close()


def test_unshard_kernel(ctx):
    import torch
    w, h, world, rb = 96, 72, 4, 8
    s, dev = _scene(ctx, dict(preset=4, algorithm="sah", segments=12))
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    rows = max(rtb200.shard_rows(rtb200.make_frame(w, h, rank=r, world=world, row_block=rb)) for r in range(world))
    gathered = torch.zeros((world, rows, w, 3), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()  # the fill above ran on torch's default stream; the handle 0 below selects the library's own stream
    stream = torch.cuda.current_stream().cuda_stream
    for rank in range(world):
        fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=rb)
        dev.render_device(s.camera, s.setting, fr, gathered[rank].data_ptr(), stream)
    image = torch.empty((h, w, 3), dtype=torch.float32, device="cuda:0")
    rtb200.unshard_device(ctx, gathered.data_ptr(), image.data_ptr(), w, h, world, rb, rows, stream)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(image.cpu().numpy()), _bits(whole))
    dev.close(); s.close()


# ---- Monte Carlo -----------------------------------------------------------------------------------
MC = [dict(preset=1, width=64, height=48, samples=4), dict(preset=2, width=64, height=48, samples=4),
      dict(preset=3, width=32, height=24, samples=2), dict(preset=2, width=40, height=30, samples=3, setting="highquality"),
      dict(preset=2, width=40, height=30, samples=3, setting="highspeed")]


@pytest.mark.parametrize("job", MC, ids=lambda j: "-".join(f"{k}{v}" for k, v in j.items()))
def test_monte_carlo_path_for_path_vs_oracle(ctx, job):
    """Same counter-based random stream on both sides: the GPU follows the same paths as the oracle.
    Differences come only from libm (acosf/cosf/sinf differ by ulps between glibc and CUDA), so nearly
    every pixel agrees to ~1e-4 relative.  (Secondary rays start exactly on the surface they left, with no
    offset, and the hemisphere is sampled uniformly, so a grazing direction that moves by one ulp can flip a
    self-intersection test; such paths diverge, which is why the bar is 97 % of pixels and 2 % of rays.)"""
    o = O.run("oracle", image=True, rng=1, seed=11, **job)
    s, dev = _scene(ctx, job)
    fr = rtb200.make_frame(job["width"], job["height"], samples=job["samples"], seed=11)
    img, st = dev.render(s.camera, _setting(s, job), fr)
    ref = o["image"]
    assert np.all(np.isfinite(img))
    close = np.abs(img - ref) <= 1e-4 * np.abs(ref) + 1e-5
    assert close.all(axis=2).mean() > 0.97, close.all(axis=2).mean()
    assert abs(st["n_rays"] - o["n_rays"]) <= 0.02 * o["n_rays"]
    assert abs(img.mean() - ref.mean()) <= 0.01 * ref.mean()
    dev.close(); s.close()


@pytest.mark.parametrize("preset,w,h", [(1, 400, 300), (2, 400, 300), (3, 120, 90)])
def test_monte_carlo_statistics_vs_erand48(ctx, preset, w, h):
    """The statistical parity criteria of SURVEY.md 8(d), as stated there, at the demo's frame size and 64 spp (preset 3
    -- 528 loose triangles per ray on the CPU -- at 120x90): the GPU (Philox, keyed by pixel and sample) returns per-pixel
    sum and sum of squares (RTB_OUTPUT_MOMENTS); the CPU side is the oracle running the REFERENCE's random stream (erand48
    seeded per row, MainWindow.cpp:273), whose image is bit-identical to the reference's (tests/test_oracle.py), with the
    same moments.  With N = spp, per colour component:
      (i)   image-mean radiance within 0.5 % -- or within 4 standard errors of the difference where the frame is too small
            for 0.5 % to be a 4-sigma event (preset 3's 120x90 frame: two CPU runs with different seeds differ by 0.7 %);
      (ii)  |mean_cpu - mean_gpu| <= 4 sqrt((s2_cpu + s2_gpu) / N) for >= 99.9 % of the pixels (all three components);
      (iii) CPU-vs-GPU RMSE <= 1.1 x CPU-vs-CPU RMSE between two disjoint seed sets (the oracle's counter stream, seeds 1 / 2).
    Measured CPU-vs-CPU in this repo's container: (i) 0.03 % / 0.24 % / 0.64 %, (ii) 99.997 % / 100 % / 100 %, (iii) 0.996-1.008."""
    spp = 64
    A = O.run("oracle", preset, width=w, height=h, samples=spp, image=True, moments=True, rng=0)
    B = O.run("oracle", preset, width=w, height=h, samples=spp, image=True, rng=1, seed=1)["image"].astype(np.float64)
    Cc = O.run("oracle", preset, width=w, height=h, samples=spp, image=True, rng=1, seed=2)["image"].astype(np.float64)
    s, dev = _scene(ctx, dict(preset=preset))
    mom, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=spp, seed=7, layout=rtb200.OUTPUT_MOMENTS))
    img, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=spp, seed=7))
    mom = mom.astype(np.float64)

    def mean_var(m):
        return m[..., :3] / spp, np.maximum(m[..., 3:] - m[..., :3] ** 2 / spp, 0.0) / (spp - 1)

    mg, vg = mean_var(mom)
    mc, vc = mean_var(A["moments"])
    assert np.allclose(mg, img, rtol=2e-5, atol=1e-6)                 # the moments describe the image that is rendered
    assert np.allclose(mc, A["image"], rtol=2e-5, atol=1e-6)
    # (i)
    diff = abs(mg.mean() - mc.mean())
    se = np.sqrt((vg.sum() + vc.sum()) / spp) / mg.size                 # standard error of the difference of the two image means
    assert diff <= max(0.005 * mc.mean(), 4.0 * se), (diff / mc.mean(), se / mc.mean())
    assert diff <= 0.01 * mc.mean()
    # (ii)
    z = np.abs(mg - mc) / np.sqrt((vg + vc) / spp + 1e-30)
    assert ((vg + vc) > 0).mean() > 0.5 or preset == 1
    within = (z <= 4.0).all(axis=2).mean()
    assert within >= 0.999, within
    # (iii)
    def rmse(a, b):
        return float(np.sqrt(np.mean((a - b) ** 2)))
    cpu_cpu = rmse(B, Cc)
    assert rmse(img.astype(np.float64), A["image"].astype(np.float64)) <= 1.1 * cpu_cpu
    assert rmse(img.astype(np.float64), B) <= 1.1 * cpu_cpu
    dev.close(); s.close()


def test_monte_carlo_is_deterministic_and_seeded(ctx):
    s, dev = _scene(ctx, dict(preset=2))
    f1 = rtb200.make_frame(48, 36, samples=8, seed=5)
    a, _ = dev.render(s.camera, s.setting, f1)
    b, _ = dev.render(s.camera, s.setting, f1)
    c, _ = dev.render(s.camera, s.setting, rtb200.make_frame(48, 36, samples=8, seed=6))
    assert np.array_equal(_bits(a), _bits(b)) and not np.array_equal(a, c)
    # sharded MC equals whole-frame MC bit for bit: the RNG key is the pixel, not the rank
    out = np.zeros_like(a)
    for rank in range(3):
        part, _ = dev.render(s.camera, s.setting, rtb200.make_frame(48, 36, samples=8, seed=5, rank=rank, world=3))
        out[rtb200.shard_row_indices(36, rank, 3, 8)] = part
    assert np.array_equal(_bits(out), _bits(a))
    dev.close(); s.close()


# ---- the reference-facing entry points ----------------------------------------------------------------
def test_script_run_dropin_path(ctx):
    """Script::Run(CudaRenderer::Render, ...) -- the RenderProc boundary of reference Scripts.h:11-12."""
    name = "p5_sah_s40_160x120"
    img, info = rtb200.script_run(5, "sah", 40, 160, 120)
    _assert_image_close(img, ARR[f"{name}.image"], name)
    assert info["n_rays"] == META[name]["n_rays"] and info["exec_ms"] >= 1
    name = "p2_simple_80x60"
    # preset 2 runs Monte Carlo through its own script; compare ray statistics only
    img, info = rtb200.script_run(2, width=80, height=60, samples=4, seed=3)
    assert np.all(np.isfinite(img)) and img.mean() > 0.05


def test_geometry_intersect_single_and_batch(ctx):
    """Geometry::intersect(Ray&) (reference Geometry.h:17) served by the GPU, one ray and batched."""
    job = dict(preset=5, algorithm="sah", segments=12, width=80, height=60)
    name = "p5_sah_s12_80x60"
    s = PresetScene(5, "sah", 12)
    # the camera rays of the golden job, rebuilt on the host exactly as Camera.cpp:20-26 does
    cam = s.camera
    w, h = 80, 60
    xs, ys = np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32))
    d = np.float32(1.0) / np.float32(h)
    sx = (xs + np.float32(0.5)) * d
    sy = np.float32(1) - (ys + np.float32(0.5)) * d
    right, up, front = (np.array(v, np.float32) for v in (cam.right, cam.up, cam.front))
    r = right[None, None, :] * ((sx - np.float32(cam.xcenter)) * np.float32(cam.fov_scale))[..., None]
    u = up[None, None, :] * ((sy - np.float32(0.5)) * np.float32(cam.fov_scale))[..., None]
    v = front[None, None, :] + r + u
    inv = np.float32(1) / np.sqrt(v[..., 0] * v[..., 0] + v[..., 1] * v[..., 1] + v[..., 2] * v[..., 2])
    dirs = v * inv[..., None]
    eye = np.array(cam.eye, np.float32)
    rays = np.concatenate([np.broadcast_to(eye + dirs * np.float32(cam.forward), dirs.shape), dirs], axis=2).reshape(-1, 6)
    hid, ht, pos, nrm = s.intersect_batch(rays)
    assert np.array_equal(hid, ARR[f"{name}.hit_id"])
    assert np.array_equal(_bits(ht), _bits(ARR[f"{name}.hit_t"]))
    for p in (0, 1234, 4799):
        i1, t1, _, _ = s.intersect_one(rays[p])
        assert i1 == hid[p] and np.float32(t1) == ht[p]
    s.close()


def test_upload_rejects_bad_scenes(ctx):
    s = PresetScene(5, "sah", 8)
    f = rtb200.FlatScene.from_buffer_copy(s.flat.contents)
    f.n_prims = 0
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    f = rtb200.FlatScene.from_buffer_copy(s.flat.contents)
    f.accel = 5  # convex: out of scope
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    # malformed accelerator / index streams are refused on the host, before anything reaches a kernel
    import ctypes as C

    def corrupted(copy_field, n, ctype, mutate):
        f = rtb200.FlatScene.from_buffer_copy(s.flat.contents)
        arr = (ctype * n)()
        C.memmove(arr, getattr(s.flat.contents, copy_field), C.sizeof(arr))
        mutate(arr)
        setattr(f, copy_field, C.cast(arr, type(getattr(f, copy_field))))
        return f, arr

    nn, nr, nt = s.flat.contents.n_kd_nodes, s.flat.contents.n_kd_refs, s.flat.contents.n_tris

    def first_inner(nodes):
        return next(i for i in range(nn) if (nodes[i].b & 3) != 3)

    def first_leaf(nodes):
        return next(i for i in range(nn) if (nodes[i].b & 3) == 3 and (nodes[i].b >> 2) > 0)

    def bad_right(nodes):  # right child not behind the left subtree
        i = first_inner(nodes); nodes[i].b = ((i + 1) << 2) | (nodes[i].b & 3)

    def right_out_of_range(nodes):
        i = first_inner(nodes); nodes[i].b = ((nn + 5) << 2) | (nodes[i].b & 3)

    def leaf_range(nodes):
        i = first_leaf(nodes); nodes[i].a = nr

    def truncated(nodes):  # the last leaf turned into an inner node: the tree never closes
        nodes[nn - 1].b = (nn << 2) | 0

    for mutate in (bad_right, right_out_of_range, leaf_range, truncated):
        f, keep = corrupted("kd_nodes", nn, rtb200.KdNode, mutate)
        with pytest.raises(rtb200.RtbError):
            ctx.upload(f)
    f, keep = corrupted("kd_leaf_tris", nr, C.c_uint32, lambda a: a.__setitem__(nr // 2, nt))
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    # a wildly out-of-range reference must be refused BEFORE the upload's own kernels index with it (k_pack_pairs):
    # the context stays usable afterwards
    f, keep = corrupted("kd_leaf_tris", nr, C.c_uint32, lambda a: a.__setitem__(0, 0x7fffffff))
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    ok = ctx.upload(s.flat)
    img, _ = ok.render(s.camera, s.setting, rtb200.make_frame(32, 24))
    assert np.isfinite(img).all()
    ok.close()
    f, keep = corrupted("tri_material", nt, C.c_int32, lambda a: a.__setitem__(nt - 1, -1))
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    f, keep = corrupted("tri_material", nt, C.c_int32, lambda a: a.__setitem__(0, 99))
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    # grid directories: word ranks must equal the running popcount of the occupancy bits, list starts must ascend
    g = PresetScene(5, "rgrid", 8)
    gf = g.flat.contents

    def grid_corrupted(field, n, ctype, mutate):
        f = rtb200.FlatScene.from_buffer_copy(gf)
        arr = (ctype * n)()
        C.memmove(arr, getattr(gf, field), C.sizeof(arr))
        mutate(arr)
        setattr(f, field, C.cast(arr, type(getattr(f, field))))
        return f, arr

    nw, nc = gf.n_cellwords, gf.n_cells_used

    def bad_rank(words):
        w = next(i for i in range(nw) if words[i].bits) + 1
        words[w].rank += 3

    def extra_bit(words):  # one more occupied cell than cell_start has lists for
        w = next(i for i in range(nw) if words[i].bits != 0xffffffff)
        words[w].bits |= (~words[w].bits) & (words[w].bits + 1)

    for mutate in (bad_rank, extra_bit):
        f, keep = grid_corrupted("grid_words", nw, rtb200.CellWord, mutate)
        with pytest.raises(rtb200.RtbError):
            ctx.upload(f)
    f, keep = grid_corrupted("grid_cell_start", nc + 1, C.c_uint32, lambda a: a.__setitem__(nc // 2, a[nc // 2 + 1] + 1))  # not ascending
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    f, keep = grid_corrupted("grid_cell_start", nc + 1, C.c_uint32, lambda a: a.__setitem__(nc, a[nc] + 5))  # past the reference array
    with pytest.raises(rtb200.RtbError):
        ctx.upload(f)
    okg = ctx.upload(g.flat)
    okg.close(); g.close()
    dev = ctx.upload(s.flat)
    with pytest.raises(rtb200.RtbError):
        dev.render(s.camera, s.setting, rtb200.make_frame(0, 10))
    with pytest.raises(rtb200.RtbError):
        dev.render(s.camera, s.setting, rtb200.make_frame(64, 48, row_block=5))
    dev.close(); s.close()


def test_4k_frame_properties(ctx):
    """BASELINE stress size, the bench's headline frame (3840x2880, preset 5, 150 segments, k-d SAH), against the CPU
    checker rendering the SAME frame: (a) hit id and hit distance of every primary ray bit-exact, (b) the Whitted image
    within 1e-5 relative on every component (the vanishing-point rows with their 21-ray chains included) and bit-exact
    on > 99.5 % of them, (c) total rays equal; then (d) the 8 row shards reassemble to the whole frame bit for bit and
    their ray counts add up."""
    w, h = 3840, 2880
    job = dict(preset=5, algorithm="sah", segments=150, width=w, height=h)
    s, dev = _scene(ctx, job)
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    assert np.all(np.isfinite(whole))
    ref = _checker(job, image=True, hits=True)
    r = dev.trace_primary(s.camera, w, h, seq=False)
    assert np.array_equal(r["hit_id"], ref["hit_id"])
    assert np.array_equal(_bits(r["hit_t"]), _bits(ref["hit_t"]))
    _assert_image_close(whole, ref["image"], "4K SAH frame")
    same = (_bits(whole).reshape(-1, 4) == _bits(ref["image"]).reshape(-1, 4)).all(axis=1).mean()
    assert same > 0.995, same
    assert st["n_rays"] == ref["n_rays"]
    # named explicitly (already covered by the whole-frame comparison): top / bottom rows, the rows and the column
    # through the vanishing point with the longest reflection chains
    for y in (0, h // 2 - 1, h // 2, h - 1):
        assert np.allclose(whole[y], ref["image"][y], rtol=RTOL, atol=1e-7)
    for x in (0, w // 2, w - 1):
        assert np.allclose(whole[:, x], ref["image"][:, x], rtol=RTOL, atol=1e-7)
    rays = 0
    for rank in range(8):
        part, pst = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, rank=rank, world=8, row_block=16))
        ys = rtb200.shard_row_indices(h, rank, 8, 16)
        assert np.array_equal(_bits(part), _bits(whole[ys]))
        rays += pst["n_rays"]
    assert rays == st["n_rays"]
    assert st["n_rays"] > 3 * w * h * 0.9  # ~3.1 rays per primary on this scene
    dev.close(); s.close()


@pytest.mark.parametrize("alg", ["rgrid"])
def test_4k_regular_grid_frame_vs_checker(ctx, alg):
    """The other headline accelerator (400 x 5 x 400 regular grid) at the bench's frame size, image and ray count."""
    w, h = 3840, 2880
    job = dict(preset=5, algorithm=alg, segments=150, width=w, height=h)
    s, dev = _scene(ctx, job)
    fr = rtb200.make_frame(w, h)
    dev.render(s.camera, s.setting, fr)
    img, st = dev.render(s.camera, s.setting, fr)  # second frame: the tiers are in play
    ref = _checker(job, image=True)
    _assert_image_close(img, ref["image"], f"4K {alg} frame")
    assert st["n_rays"] == ref["n_rays"]
    dev.close(); s.close()


# ---- edge cases ----------------------------------------------------------------------------------------
RAGGED = [dict(preset=5, algorithm="sah", segments=12, width=37, height=23),
          dict(preset=5, algorithm="rgrid", segments=12, width=9, height=5),
          dict(preset=4, algorithm="fgrid", segments=8, width=1, height=1),
          dict(preset=4, algorithm="kd", segments=12, width=101, height=3),
          dict(preset=2, width=13, height=7, setting="simple")]


@pytest.mark.parametrize("job", RAGGED, ids=lambda j: "-".join(f"{k}{v}" for k, v in j.items()))
def test_ragged_frame_sizes(ctx, job):
    """Widths that are not a multiple of the 8x4 tile, single-pixel and single-row frames; each frame is
    rendered twice so the second pass runs with the heaviest-first tile order learnt from the first."""
    o = O.run("oracle", image=True, hits=True, seq=True, **job)
    s, dev = _scene(ctx, job)
    r = dev.trace_primary(s.camera, job["width"], job["height"], seq=True)
    assert np.array_equal(r["hit_id"], o["hit_id"]) and np.array_equal(r["seq_hash"], o["seq_hash"])
    fr = rtb200.make_frame(job["width"], job["height"])
    a, st = dev.render(s.camera, _setting(s, job), fr)
    b, _ = dev.render(s.camera, _setting(s, job), fr)
    _assert_image_close(a, o["image"], str(job))
    assert np.array_equal(_bits(a), _bits(b)) and st["n_rays"] == o["n_rays"]
    dev.close(); s.close()


def test_more_ranks_than_row_blocks(ctx):
    """world larger than the number of row blocks: some ranks own nothing and render nothing."""
    s, dev = _scene(ctx, dict(preset=5, algorithm="sah", segments=12))
    w, h, world = 64, 20, 8  # 3 blocks of 8 rows
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    out, rays = np.zeros_like(whole), 0
    for rank in range(world):
        fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8)
        part, pst = dev.render(s.camera, s.setting, fr)
        ys = rtb200.shard_row_indices(h, rank, world, 8)
        assert part.shape[0] == len(ys)
        if len(ys):
            out[ys] = part
            rays += pst["n_rays"]
    assert np.array_equal(_bits(out), _bits(whole)) and rays == st["n_rays"]
    dev.close(); s.close()


@pytest.mark.parametrize("max_depth", [0, 1, 2, 7])
def test_max_depth_settings(ctx, max_depth):
    """RenderSetting.maxDepth cuts the reflection chain exactly where the reference's `++depth > maxDepth`
    does (MainWindow.cpp:83); the oracle has no knob for it, so check monotone ray counts and the limits."""
    s, dev = _scene(ctx, dict(preset=5, algorithm="sah", segments=12))
    st_ = rtb200.RenderSetting(0, max_depth, rtb200.INT_MAX, 0)
    img, st = dev.render(s.camera, st_, rtb200.make_frame(80, 60))
    n = 80 * 60
    if max_depth == 0:
        assert st["n_rays"] == n and not img.any()  # every first hit already exceeds the depth: black
    else:
        full, fst = dev.render(s.camera, s.setting, rtb200.make_frame(80, 60))
        assert n <= st["n_rays"] <= fst["n_rays"]
        assert st["n_rays"] <= n * (max_depth + 1)  # the ray that exceeds the depth is still traced (line 71-84)
    dev.close(); s.close()


def test_tile_order_does_not_change_results(ctx):
    """Heaviest-first scheduling and the resumable traversal of the latency-critical tiles only change who
    computes a pixel when: ten consecutive frames are bit-identical and so is a fresh context's first frame."""
    job = dict(preset=5, algorithm="sah", segments=40)
    s, dev = _scene(ctx, job)
    fr = rtb200.make_frame(320, 240)
    first, st0 = dev.render(s.camera, s.setting, fr)
    for _ in range(9):
        again, st = dev.render(s.camera, s.setting, fr)
        assert np.array_equal(_bits(first), _bits(again)) and st["n_rays"] == st0["n_rays"]
    ctx2 = rtb200.Context(0)
    dev2 = ctx2.upload(s.flat)
    fresh, _ = dev2.render(s.camera, s.setting, fr)
    assert np.array_equal(_bits(first), _bits(fresh))
    dev2.close(); ctx2.close(); dev.close(); s.close()


@pytest.mark.parametrize("alg", ["sah", "kd", "rgrid", "fgrid"])
def test_three_kernel_tiers_agree(ctx, alg):
    """Once a tile order exists a frame is shared by three kernels (warp-per-pixel for the heaviest tiles,
    resumable walk for the latency-critical ones, per-ray walk for the rest).  Frames 2..4 must reproduce
    frame 1 (rendered by one kernel in raster order) bit for bit, with identical ray / triangle-test /
    traversal-step totals -- and those totals are the oracle's."""
    job = dict(preset=5, algorithm=alg, segments=40, width=256, height=192)
    o = O.run("oracle", image=True, **job)
    s, dev = _scene(ctx, job)
    fr = rtb200.make_frame(256, 192, counters=1)
    first, st0 = dev.render(s.camera, s.setting, fr)
    assert (st0["n_rays"], st0["n_tri_tests"], st0["n_steps"]) == (o["n_rays"], o["n_tri_tests"], o["n_steps"])
    for _ in range(3):
        again, st = dev.render(s.camera, s.setting, fr)
        assert np.array_equal(_bits(first), _bits(again))
        assert (st["n_rays"], st["n_tri_tests"], st["n_steps"]) == (st0["n_rays"], st0["n_tri_tests"], st0["n_steps"])
        assert st["n_launches"] >= 5  # at least two render kernels + the three tile-order kernels
    dev.close(); s.close()


def test_moving_camera_keeps_results_exact(ctx):
    """A camera move reuses the previous frame's tile order for one frame (without the latency tiers); the image of
    every frame equals what a fresh context renders for that camera."""
    job = dict(preset=5, algorithm="sah", segments=40)
    s, dev = _scene(ctx, job)
    fr = rtb200.make_frame(256, 192)
    for step in range(4):
        cam = type(s.camera).from_buffer_copy(s.camera)
        cam.eye[0] += 0.75 * step
        cam.eye[2] -= 1.5 * step
        ctx2 = rtb200.Context(0)  # no history: raster order, one kernel
        dev2 = ctx2.upload(s.flat)
        ref, st2 = dev2.render(cam, s.setting, fr)
        dev2.close(); ctx2.close()
        for _ in range(2):  # moved frame (order only), then the same view again (tiers)
            img, st = dev.render(cam, s.setting, fr)
            assert np.array_equal(_bits(img), _bits(ref)) and st["n_rays"] == st2["n_rays"]
    dev.close(); s.close()


# ---- GPU-side grid builder ----------------------------------------------------------
@pytest.mark.parametrize("job", [dict(preset=5, algorithm="rgrid", segments=150), dict(preset=5, algorithm="fgrid", segments=150),
                                 dict(preset=4, algorithm="rgrid", segments=40), dict(preset=4, algorithm="fgrid", segments=12)],
                         ids=lambda j: f"p{j['preset']}_{j['algorithm']}_s{j['segments']}")
def test_grid_built_on_device_is_identical(ctx, job):
    """rtb_scene_upload without grid arrays (grid_build_resolution = 400) builds the grid on the device: dims, occupied
    cells, references, longest list and the canonical structure hash equal the host builder's (= the reference's,
    tests/test_host_builders.py), and the rendered frame is bit-identical."""
    host = PresetScene(job["preset"], job["algorithm"], job["segments"])
    rtb200.set_grid_on_device(True)
    try:
        lazy = PresetScene(job["preset"], job["algorithm"], job["segments"])
    finally:
        rtb200.set_grid_on_device(False)
    assert lazy.flat.contents.grid_build_resolution == 400 and not lazy.flat.contents.grid_words
    dh, dl = ctx.upload(host.flat), ctx.upload(lazy.flat)
    hh, sh = dh.grid_hash()
    hl, sl = dl.grid_hash()
    st = host.stats()
    assert hl == hh == host.struct_hash()
    assert sl == sh and {k: st[k] for k in sl} == sl
    fr = rtb200.make_frame(160, 120)
    a, sa = dh.render(host.camera, host.setting, fr)
    b, sb = dl.render(lazy.camera, lazy.setting, fr)
    assert np.array_equal(_bits(a), _bits(b)) and sa["n_rays"] == sb["n_rays"]
    dh.close(); dl.close(); host.close(); lazy.close()


@pytest.mark.parametrize("preset,segments", [(5, 12), (5, 40), (4, 24), (5, 150), (4, 150)])
def test_sah_tree_built_on_device_is_identical(ctx, preset, segments):
    """SURVEY 8f rank 2, second half: rtb_scene_upload with RTB_ACCEL_KD_SAH, no node arrays and build parameters builds the
    tree on the device (csrc/rtb_build_kd.cuh: reference Tunnel.cpp:546-638, 671-784 -- candidate evaluation as histogram +
    prefix sums per node, level-synchronous).  The node array and the leaf reference array read back from the device are
    the host builder's, ELEMENT FOR ELEMENT (the host builder's tree is the reference's: tests/test_host_builders.py, leaf
    counts 34478 / 52778 of the reference's logs at 150 segments), and so are the root box and the rendered frame."""
    host = PresetScene(preset, "sah", segments)
    rtb200.set_kd_on_device(True)
    try:
        lazy = PresetScene(preset, "sah", segments)
    finally:
        rtb200.set_kd_on_device(False)
    hf, lf = host.flat.contents, lazy.flat.contents
    assert lf.kd_build_max_depth == 18 and lf.kd_build_leaf_size == 8 and lf.kd_build_candidates == 100 and not lf.kd_nodes
    dh, dl = ctx.upload(host.flat), ctx.upload(lazy.flat)
    nodes, refs, levels, box = dl.kd_download()
    want_nodes = np.ctypeslib.as_array(C.cast(hf.kd_nodes, C.POINTER(C.c_uint32)), shape=(hf.n_kd_nodes, 2))
    want_refs = np.ctypeslib.as_array(hf.kd_leaf_tris, shape=(hf.n_kd_refs,))
    assert nodes.shape == want_nodes.shape and np.array_equal(nodes, want_nodes)
    assert np.array_equal(refs, want_refs)
    assert 1 <= levels <= 20
    assert np.array_equal(box[:3], np.array(list(hf.kd_min), np.float32))
    assert np.array_equal(box[3:], np.array(list(hf.kd_max), np.float32) - np.array(list(hf.kd_min), np.float32))
    assert dl.upload_bytes < dh.upload_bytes  # nothing of the accelerator crosses PCIe
    fr = rtb200.make_frame(160, 120, counters=1)
    a, sa = dh.render(host.camera, host.setting, fr)
    b, sb = dl.render(lazy.camera, lazy.setting, fr)
    assert np.array_equal(_bits(a), _bits(b)) and (sa["n_rays"], sa["n_tri_tests"], sa["n_steps"]) == (sb["n_rays"], sb["n_tri_tests"], sb["n_steps"])
    dh.close(); dl.close(); host.close(); lazy.close()


@pytest.mark.parametrize("name", ["p5_rgrid_s24_96x72", "p4_rgrid_s16_64x48", "p5_rgrid_s150_400x300"])
def test_exact_grid_binning_host_and_device(ctx, name):
    """the exact triangle / cell overlap test (reference Triangle.cpp:152-199, compiled out at
    Tunnel.cpp:435-445) as an option of BOTH grid builders.  Fixtures come from the reference rebuilt with that branch
    enabled (libref_sat.so): the host-built and the device-built grid have its structure hash and statistics, and the
    frames rendered over them have its hit ids, distances, cell sequences, counts and image."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "sat_golden.json")) as f:
        g = json.load(f)[name]
    sat = np.load(os.path.join(os.path.dirname(__file__), "golden", "sat_golden.npz"))
    job = g["job"]
    rtb200.set_exact_grid_binning(True)
    try:
        host = PresetScene(job["preset"], job["algorithm"], job["segments"])
        rtb200.set_grid_on_device(True)
        lazy = PresetScene(job["preset"], job["algorithm"], job["segments"])
    finally:
        rtb200.set_grid_on_device(False)
        rtb200.set_exact_grid_binning(False)
    assert lazy.flat.contents.grid_build_exact == 1 and not lazy.flat.contents.grid_words
    import hashlib
    for s in (host, lazy):
        dev = ctx.upload(s.flat)
        h, st = dev.grid_hash()
        assert f"{h:016x}" == g["struct_hash"]
        assert {k: g["stats"][k] for k in st} == st
        r = dev.trace_primary(s.camera, job["width"], job["height"], seq=True)
        for k in ("hit_id", "hit_t", "seq_len", "seq_hash"):
            assert hashlib.sha256(np.ascontiguousarray(r[k]).tobytes()).hexdigest() == g["sha256"][k], k
        fr = rtb200.make_frame(job["width"], job["height"], counters=1)
        img, cnt = dev.render(s.camera, s.setting, fr)
        img2, _ = dev.render(s.camera, s.setting, fr)  # tiers
        assert np.array_equal(_bits(img), _bits(img2))
        assert (cnt["n_rays"], cnt["n_tri_tests"], cnt["n_steps"]) == (g["n_rays"], g["n_tri_tests"], g["n_steps"])
        if f"{name}.image" in sat.files:
            _assert_image_close(img, sat[f"{name}.image"], name)
        dev.close(); s.close()


# ---- output stage: saturate -> 8 bit, BMP writer -----------------------------------
@pytest.mark.parametrize("name", ["p4_sah_s12_80x60", "p5_rgrid_s40_160x120", "p1_simple_80x60", "p2_simple_80x60"])
def test_rgb8_output_stage_vs_reference(ctx, name):
    """RTB_OUTPUT_RGB8 runs the reference's output stage on the GPU (MainWindow.cpp:305-311).  A byte can
    differ by one only where the float differs in the last ulps across a quantisation step (Phong powf)."""
    job = META[name]["job"]
    s, dev = _scene(ctx, job)
    fr = rtb200.make_frame(job["width"], job["height"], layout=rtb200.OUTPUT_RGB8)
    img8, _ = dev.render(s.camera, _setting(s, job), fr)
    ref8 = ARR[f"{name}.image8"]
    assert img8.dtype == np.uint8 and img8.shape == ref8.shape
    diff = np.abs(img8.astype(np.int16) - ref8.astype(np.int16))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3
    # and it is exactly the quantisation of this library's own float image
    f32, _ = dev.render(s.camera, _setting(s, job), rtb200.make_frame(job["width"], job["height"]))
    q = (np.minimum(f32, np.float32(1.0)) * np.float32(255)).astype(np.int32).astype(np.uint8)
    assert np.array_equal(q, img8)
    dev.close(); s.close()


def test_dropin_8bit_and_bitmap(ctx, tmp_path):
    name = "p5_sah_s40_160x120"
    bmp = str(tmp_path / "frame.bmp")
    img8, info = rtb200.script_run8(5, "sah", 40, 160, 120, bmp_path=bmp)
    assert np.array_equal(img8, ARR[f"{name}.image8"]) and info["n_rays"] == META[name]["n_rays"]
    raw = open(bmp, "rb").read()
    assert raw[:2] == b"BM" and len(raw) == 54 + 160 * 120 * 3
    w, h, bpp = np.frombuffer(raw[18:26], np.int32)[0], np.frombuffer(raw[18:26], np.int32)[1], np.frombuffer(raw[28:30], np.int16)[0]
    assert (w, h, bpp) == (160, 120, 24)
    rows = np.frombuffer(raw[54:], np.uint8).reshape(120, 160, 3)[::-1, :, ::-1]  # bottom-up BGR -> top-down RGB
    assert np.array_equal(rows, img8)


# ---- PerformanceTest console benchmark and convex accelerators (rank 3) --------------
@pytest.mark.parametrize("alg", ["convex", "convexsimple"])
def test_convex_accelerators_vs_oracle(ctx, alg):
    """Convex / ConvexSimple (PerformanceTest/ConvexAcc.cpp) on 4,000 rays of the full-size 150 x 150 tunnel, bit-exact
    against the oracle (itself bit-exact against the compiled reference): reached, depth, last hit id / position."""
    xy = np.random.default_rng(17).random((4000, 2), dtype=np.float32)
    o = O.bounce("oracle", xy, 2000.0, 1.5708, 150, 150, alg, pt_builders=True)
    g = rtb200.perf_test(xy, 2000.0, 1.5708, 150, 150, alg)
    for k in ("reached", "depth", "last_id"):
        assert np.array_equal(g[k], o[k]), (k, int((g[k] != o[k]).sum()))
    assert np.array_equal(_bits(g["last_pos"]), _bits(o["last_pos"]))
    assert g["total_rays"] == o["total_rays"]


@pytest.mark.parametrize("name", ["p5_convex_s40_160x120", "p5_convexsimple_s40_160x120", "p4_convex_s24_160x120"])
def test_convex_presets_vs_reference(ctx, name):
    """RayTracingOpt's own convex walk behind the preset scenes (tests/golden/convex_golden.*, recorded from libref.so):
    primary hit ids and distances bit-exact, Whitted image within 1e-5 (bit-exact where no Phong highlight is seen), ray and
    triangle-test totals equal."""
    with open(os.path.join(os.path.dirname(__file__), "golden", "convex_golden.json")) as f:
        m = json.load(f)[name]
    arr = np.load(os.path.join(os.path.dirname(__file__), "golden", "convex_golden.npz"))
    job = m["job"]
    s, dev = _scene(ctx, job)
    r = dev.trace_primary(s.camera, job["width"], job["height"], seq=False)
    assert np.array_equal(r["hit_id"], arr[f"{name}.hit_id"])
    assert np.array_equal(_bits(r["hit_t"]), _bits(arr[f"{name}.hit_t"]))
    for _ in range(2):  # raster-order frame, then the ordered one
        img, st = dev.render(s.camera, s.setting, rtb200.make_frame(job["width"], job["height"], counters=1))
        _assert_image_close(img, arr[f"{name}.image"], name)
        assert (st["n_rays"], st["n_tri_tests"]) == (m["n_rays"], m["n_tri_tests"])
    if job["preset"] == 5:
        same = (_bits(img).reshape(-1, 4) == _bits(arr[f"{name}.image"]).reshape(-1, 4)).all(axis=1).mean()
        assert same > 0.995, same
    dev.close(); s.close()


def test_convex_accelerator_renders_and_single_rays(ctx):
    """The convex accelerator behind the other entry points: a Whitted frame (the context travels down the reflection
    chain) equals the per-ray API applied bounce by bounce is not required -- here: the frame is deterministic, finite,
    and primary hits agree with the SAH tree wherever both report the nearest triangle (the convex walk returns the
    first accepted triangle of the wall segment, which for a ray inside a convex segment is the only one)."""
    sc = rtb200.PerfScene(2000.0, 1.5708, 60, 60, "convex")
    ss = rtb200.PerfScene(2000.0, 1.5708, 60, 60, "sah")
    dc, ds = ctx.upload(sc.flat), ctx.upload(ss.flat)
    a = dc.trace_primary(sc.camera, 160, 120)
    b = ds.trace_primary(ss.camera, 160, 120)
    agree = (a["hit_id"] == b["hit_id"]).mean()
    assert agree > 0.999, agree
    img1, st1 = dc.render(sc.camera, sc.setting, rtb200.make_frame(160, 120))
    img2, st2 = dc.render(sc.camera, sc.setting, rtb200.make_frame(160, 120))
    assert np.isfinite(img1).all() and np.array_equal(_bits(img1), _bits(img2)) and st1["n_rays"] == st2["n_rays"]
    dc.close(); ds.close(); sc.close(); ss.close()


# ---- PerformanceTest console benchmark ---------------------------------------------
@pytest.mark.parametrize("alg", ["rgrid", "fgrid", "kd", "sah", "linear"])
def test_performance_test_bounce_workload(ctx, alg):
    """src/PerformanceTest/main.cpp: random camera rays mirror-bounced through the tunnel until they hit
    the plane at its exit (<= 200 reflections).  Bit-exact against the oracle (itself bit-exact against
    the reference's classes, tests/test_oracle.py): reached flag, depth, last hit id and position per
    ray; and the reference's own assertion -- no ray may miss (main.cpp:158-161)."""
    rng = np.random.default_rng(11)
    n, seg = (64, 10) if alg == "linear" else (500, 30)
    xy = rng.random((n, 2), dtype=np.float32)
    o = O.bounce("oracle", xy, 2000.0, 1.5708, seg, seg, alg, pt_builders=True)  # that program's own generator / builders
    g = rtb200.perf_test(xy, 2000.0, 1.5708, seg, seg, alg)
    for k in ("reached", "depth", "last_id"):
        assert np.array_equal(g[k], o[k]), k
    assert np.array_equal(_bits(g["last_pos"]), _bits(o["last_pos"]))
    assert g["total_rays"] == o["total_rays"] and g["reached"].all()


@pytest.mark.parametrize("alg", ["rgrid", "sah", "kd"])
def test_performance_test_large_batch_one_lane_per_ray(ctx, alg):
    """Batches beyond two warps per resident warp slot (18,944 rays) take the one-lane-per-ray kernel, smaller ones the
    warp-per-ray kernel (k_bounce_rays<WIDE>): both must reproduce the oracle bit for bit, and agree with each other on
    the rays they share."""
    xy = np.random.default_rng(12).random((24000, 2), dtype=np.float32)
    o = O.bounce("oracle", xy, 2000.0, 1.5708, 30, 30, alg, pt_builders=True)
    big = rtb200.perf_test(xy, 2000.0, 1.5708, 30, 30, alg)
    small = rtb200.perf_test(xy[:700], 2000.0, 1.5708, 30, 30, alg)
    for k in ("reached", "depth", "last_id"):
        assert np.array_equal(big[k], o[k]), k
        assert np.array_equal(small[k], o[k][:700]), k
    assert np.array_equal(_bits(big["last_pos"]), _bits(o["last_pos"]))
    assert np.array_equal(_bits(small["last_pos"]), _bits(o["last_pos"][:700]))
    assert big["total_rays"] == o["total_rays"]


@pytest.mark.parametrize("case", ["r2000_s30", "r100_a75_s24x12"])
def test_performance_test_program_golden(ctx, case):
    """The same workload against fixtures recorded from the compiled PerformanceTest sources themselves
    (oracle/_ref/libref_pt.so -> tests/golden/bounce_pt_golden.npz): grids, median tree, event-sweep SAH tree and the
    convex accelerators (first accepted triangle of the wall segment, ray context carried over bounces)."""
    cases = {"r2000_s30": (2000.0, 1.5708, 30, 30), "r100_a75_s24x12": (100.0, 1.309, 24, 12)}
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "bounce_pt_golden.npz"))
    for alg in ("rgrid", "kd", "sah", "convex", "convexsimple"):
        g = rtb200.perf_test(gold["xy"], *cases[case], alg)
        for k in ("reached", "depth", "last_id", "last_pos"):
            assert np.array_equal(_bits(g[k]), _bits(gold[f"{case}.{alg}.{k}"])), (alg, k)


@pytest.mark.gpu
def test_packed_pretest_equals_scalar(ctx):
    """The kernels decide two triangles per call on the packed FP32 pipe (rtb_pretest.h: sureReject2); each half must
    take, bit for bit, the decision of the scalar sureReject that tests/test_pretest.py checks against the oracle."""
    import ctypes as C
    lib = rtb200.cuda_lib()
    lib.rtb_selftest_pretest.argtypes = [C.c_void_p, C.c_int64, C.c_uint64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    mism, rej = C.c_int64(-1), C.c_int64(-1)
    n = 20_000_000
    rc = lib.rtb_selftest_pretest(ctx._h, n, 20261018, C.byref(mism), C.byref(rej))
    assert rc == 0
    assert mism.value == 0
    assert 0.05 * n < rej.value < 0.95 * n  # both outcomes are exercised
