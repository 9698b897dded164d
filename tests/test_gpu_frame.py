"""GPU tests of the frame-level entry points added with ABI 4 (include/rtb.h): one frame buffer filled by several
shards / devices (RTB_LAYOUT_GLOBAL, rtb_multi_*, rtb_ipc_*), the Monte-Carlo accumulation buffer and sample shards
(RTB_OUTPUT_MOMENTS, rtb_frame.sample_first / sample_count), progress / log reporting behind the reference's
RenderProc (Scripts.h:9-12), and the first-frame schedule (rtb_forget_schedule).  Everything goes through the C ABI;
the reference values are whole frames rendered by the same library, whose parity with the reference
tests/test_gpu_parity.py establishes."""
import os
import subprocess
import sys

import numpy as np
import pytest

import rtb200
from rtb200 import PresetScene

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    c = rtb200.Context(0)
    yield c
    c.close()


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def _n_gpus():
    import torch
    return torch.cuda.device_count()


# ---- RTB_LAYOUT_GLOBAL: shards store straight into the whole frame -------------------------------------------
@pytest.mark.parametrize("world,col_block,alg", [(2, 0, "sah"), (3, 0, "rgrid"), (4, 16, "sah"), (8, 8, "kd"), (8, 32, "fgrid")])
def test_global_layout_shards_fill_one_frame(ctx, world, col_block, alg):
    """Every rank of a frame is given the SAME page-locked frame and stores only its own pixels, at y * width + x: after
    the last rank the buffer equals the whole-frame render bit for bit (float and 8-bit), with no gather / unshard pass.
    Rendered twice: the second pass runs through the tile order and the latency tiers."""
    w, h = 256, 96
    s = PresetScene(5, alg, 24)
    dev = ctx.upload(s.flat)
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    whole8, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, layout=rtb200.OUTPUT_RGB8))
    frame = rtb200.PinnedArray((h, w, 3))
    frame8 = rtb200.PinnedArray(((h * w * 3 + 3) // 4,))
    bytes8 = frame8.array.view(np.uint8)[: h * w * 3].reshape(h, w, 3)
    for repeat in range(2):
        frame.array[:] = -1.0
        bytes8[:] = 7
        rays = 0
        for rank in range(world):
            fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8, col_block=col_block, layout=rtb200.LAYOUT_GLOBAL)
            _, pst = dev.render(s.camera, s.setting, fr, out=frame.array)
            rays += pst["n_rays"]
            fr8 = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8, col_block=col_block,
                                    layout=rtb200.LAYOUT_GLOBAL | rtb200.OUTPUT_RGB8)
            dev.render(s.camera, s.setting, fr8, out=bytes8)
        assert np.array_equal(_bits(frame.array), _bits(whole))
        assert np.array_equal(bytes8, whole8)
        assert rays == st["n_rays"]
    # a pageable whole-frame buffer cannot take shard stores: refused, not silently mis-assembled
    with pytest.raises(rtb200.RtbError):
        dev.render(s.camera, s.setting, rtb200.make_frame(w, h, rank=0, world=2, layout=rtb200.LAYOUT_GLOBAL), out=np.zeros((h, w, 3), np.float32))
    frame.close(); frame8.close(); dev.close(); s.close()


def test_global_layout_device_frame(ctx):
    """The same through rtb_render_device into ONE device frame (what the ranks of a torch.distributed job do with the
    owner's IPC-mapped frame)."""
    import torch
    w, h, world = 192, 144, 4
    s = PresetScene(5, "sah", 24)
    dev = ctx.upload(s.flat)
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    tstream = torch.cuda.Stream()  # a real stream handle (the default stream's handle 0 would select the library's own stream)
    with torch.cuda.stream(tstream):
        image = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda:0")
        for col_block in (0, 8):
            image.fill_(-1.0)
            for rank in range(world):
                fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8, col_block=col_block, layout=rtb200.LAYOUT_GLOBAL)
                dev.render_device(s.camera, s.setting, fr, image.data_ptr(), tstream.cuda_stream)
            tstream.synchronize()
            assert np.array_equal(_bits(image.cpu().numpy()), _bits(whole))
    dev.close(); s.close()


@pytest.mark.parametrize("world,col_block", [(2, 0), (4, 0), (8, 32), (4, 16)])
def test_scatter_shard_into_whole_frame(ctx, world, col_block):
    """rtb_scatter_shard_device: a rank's LOCAL shard image moved to its place in the whole frame by one streaming copy --
    the alternative to RTB_LAYOUT_GLOBAL's stores from inside the render kernels; all ranks together give the whole frame."""
    import torch
    w, h = 512, 96
    s = PresetScene(5, "sah", 24)
    dev = ctx.upload(s.flat)
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    tstream = torch.cuda.Stream()  # a real stream handle (the default stream's handle 0 would select the library's own stream)
    with torch.cuda.stream(tstream):
        image = torch.full((h, w, 3), -1.0, dtype=torch.float32, device="cuda:0")
        for rank in range(world):
            fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8, col_block=col_block)
            local = torch.zeros((rtb200.shard_rows(fr), rtb200.shard_width(fr), 3), dtype=torch.float32, device="cuda:0")
            dev.render_device(s.camera, s.setting, fr, local.data_ptr(), tstream.cuda_stream)
            ctx.scatter_shard_device(local.data_ptr(), image.data_ptr(), fr, tstream.cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(_bits(image.cpu().numpy()), _bits(whole))
    dev.close(); s.close()


def test_global_layout_monte_carlo_and_reference_order(ctx):
    w, h, world = 64, 48, 3
    s = PresetScene(2)
    dev = ctx.upload(s.flat)
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=3, seed=4))
    frame = rtb200.PinnedArray((h, w, 3))
    for rank in range(world):
        dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=3, seed=4, rank=rank, world=world, layout=rtb200.LAYOUT_GLOBAL),
                   out=frame.array)
    assert np.array_equal(_bits(frame.array), _bits(whole))
    # the reference framebuffer order (index = x * height + y, MainWindow.cpp:276) addresses by frame coordinates: with
    # RTB_LAYOUT_GLOBAL row shards fill one such buffer
    col = rtb200.PinnedArray((w, h, 3))
    for rank in range(world):
        fr = rtb200.make_frame(w, h, samples=3, seed=4, rank=rank, world=world, layout=rtb200.LAYOUT_GLOBAL | rtb200.LAYOUT_REFERENCE)
        dev.render(s.camera, s.setting, fr, out=col.array)
    assert np.array_equal(_bits(col.array.transpose(1, 0, 2)), _bits(whole))
    frame.close(); col.close(); dev.close(); s.close()


# ---- rtb_multi_*: one host thread, n devices, one assembled host frame ------------------------------------------
@pytest.mark.parametrize("devices", [[0], [0, 0], [0, 0, 0, 0, 0]])
def test_multi_render_assembles_one_host_frame(ctx, devices):
    """rtb_multi_render over n contexts (here: n contexts on device 0, which exercises the same sharding, launch and
    assembly code as n GPUs; the all-GPU variant follows) returns the whole frame in ONE host buffer -- page-locked or
    pageable -- identical to rtb_render's, counts included; 5 contexts take the column-block path."""
    w, h = 320, 120
    s = PresetScene(5, "sah", 24)
    dev = ctx.upload(s.flat)
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, counters=1))
    whole8, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, layout=rtb200.OUTPUT_RGB8))
    m = rtb200.MultiContext(len(devices), devices)
    ms = m.upload(s.flat)
    assert ms.upload_bytes == len(devices) * dev.upload_bytes
    pinned = rtb200.PinnedArray((h, w, 3))
    for repeat in range(3):
        pinned.array[:] = -1.0
        _, mst = ms.render(s.camera, s.setting, rtb200.make_frame(w, h, counters=1), out=pinned.array)
        assert np.array_equal(_bits(pinned.array), _bits(whole))
        assert (mst["n_rays"], mst["n_tri_tests"], mst["n_steps"]) == (st["n_rays"], st["n_tri_tests"], st["n_steps"])
    pageable, _ = ms.render(s.camera, s.setting, rtb200.make_frame(w, h))
    assert np.array_equal(_bits(pageable), _bits(whole))
    out8, _ = ms.render(s.camera, s.setting, rtb200.make_frame(w, h, layout=rtb200.OUTPUT_RGB8))
    assert np.array_equal(out8, whole8)
    refo, _ = ms.render(s.camera, s.setting, rtb200.make_frame(w, h, layout=rtb200.LAYOUT_REFERENCE))
    assert refo.shape == (w, h, 3) and np.array_equal(_bits(refo.transpose(1, 0, 2)), _bits(whole))
    pinned.close(); ms.close(); m.close(); dev.close(); s.close()


@pytest.mark.parametrize("algorithm,size", [("sah", (640, 480)), ("rgrid", (640, 480)), ("fgrid", (320, 240)), ("sah", (328, 200))])
def test_host_frames_leave_in_groups_of_four_tiles(ctx, algorithm, size):
    """Frames rendered straight into page-locked host memory: a CTA of the throughput kernels takes four adjacent tiles and
    stores 384-byte row segments (FrameParams::group4, storeGroup), the tile order is an order of such groups.  Pixels, counters
    and the 8-bit output are those of the pageable path (device frame + copy, tile-granular order) frame after frame -- raster
    first frame, learnt order, tiers -- also when frames of both kinds of the same view alternate on one context (the order a
    tile-granular frame leaves behind is no group order), and for a width that is no multiple of 32 (no group mode)."""
    w, h = size
    s = PresetScene(5, algorithm, 40)
    dev = ctx.upload(s.flat)
    fr = rtb200.make_frame(w, h, counters=1)
    ref, st = dev.render(s.camera, s.setting, fr)                       # pageable buffer: device frame + copy
    ref8, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, layout=rtb200.OUTPUT_RGB8))
    pinned = rtb200.PinnedArray((h, w, 3))
    pinned8 = rtb200.PinnedArray(((h * w * 3 + 3) // 4,))
    host8 = pinned8.array.view(np.uint8)[: h * w * 3].reshape(h, w, 3)
    for i in range(5):
        pinned.array[:] = -1.0
        _, st2 = dev.render(s.camera, s.setting, fr, out=pinned.array)
        assert np.array_equal(_bits(pinned.array), _bits(ref)), f"frame {i}"
        assert (st2["n_rays"], st2["n_tri_tests"], st2["n_steps"]) == (st["n_rays"], st["n_tri_tests"], st["n_steps"])
        if i == 2:  # a tile-granular frame of the same view in between
            again, _ = dev.render(s.camera, s.setting, fr)
            assert np.array_equal(_bits(again), _bits(ref))
    for i in range(3):
        host8[:] = 7
        dev.render(s.camera, s.setting, rtb200.make_frame(w, h, layout=rtb200.OUTPUT_RGB8), out=host8)
        assert np.array_equal(host8, ref8)
    # the reference's column-major order (the RenderProc drop-in's): blocks of 8x16 pixels, 192-byte column segments
    col = pinned.array.reshape(w, h, 3)
    col8 = host8.reshape(w, h, 3)
    for i in range(4):
        col[:] = -1.0
        _, st3 = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, counters=1, layout=rtb200.LAYOUT_REFERENCE), out=col)
        assert np.array_equal(_bits(col.transpose(1, 0, 2)), _bits(ref)), f"reference-order frame {i}"
        assert st3["n_rays"] == st["n_rays"]
        col8[:] = 9
        dev.render(s.camera, s.setting, rtb200.make_frame(w, h, layout=rtb200.LAYOUT_REFERENCE | rtb200.OUTPUT_RGB8), out=col8)
        assert np.array_equal(col8.transpose(1, 0, 2), ref8)
    pinned.close(); pinned8.close(); dev.close(); s.close()


@pytest.mark.parametrize("algorithm", ["sah", "rgrid"])
def test_upload_out_of_page_locked_scene_arrays(ctx, algorithm):
    """rtb_flat_scene.arrays_page_locked: the upload copies H2D straight out of the caller's (registered) arrays instead of
    staging them -- same device scene, same frame, same byte accounting; one context and a device set (where the first
    device's upload publishes the source addresses to the replicas)."""
    w, h = 256, 96
    s = PresetScene(5, algorithm, 40)
    dev = ctx.upload(s.flat)
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, counters=1))
    staged_bytes = dev.upload_bytes
    dev.close()
    s.pin()
    assert s.flat.contents.arrays_page_locked == 1 and len(s._pinned) >= 2
    for _ in range(3):  # repeated uploads out of the same arrays
        dev = ctx.upload(s.flat)
        img, st2 = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, counters=1))
        assert dev.upload_bytes == staged_bytes
        assert np.array_equal(_bits(img), _bits(whole)) and st2["n_rays"] == st["n_rays"] and st2["n_tri_tests"] == st["n_tri_tests"]
        dev.close()
    m = rtb200.MultiContext(3, [0, 0, 0])
    for _ in range(2):
        ms = m.upload(s.flat)
        assert ms.upload_bytes == 3 * staged_bytes
        out, mst = ms.render(s.camera, s.setting, rtb200.make_frame(w, h, counters=1))
        assert np.array_equal(_bits(out), _bits(whole)) and mst["n_rays"] == st["n_rays"]
        ms.close()
    m.close()
    s.unpin()
    assert s.flat.contents.arrays_page_locked == 0
    dev = ctx.upload(s.flat)  # and staged again
    img, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    assert np.array_equal(_bits(img), _bits(whole))
    dev.close(); s.close()


def test_multi_upload_reports_a_bad_scene_once(ctx):
    """A flat scene the first device's checks refuse: the replicas (which wait for that verdict) are released, the call fails with
    the first device's message and the device set stays usable."""
    s = PresetScene(5, "sah", 24)
    f = s.flat.contents
    refs = np.ctypeslib.as_array(f.kd_leaf_tris, shape=(f.n_kd_refs,))
    keep = int(refs[5])
    m = rtb200.MultiContext(4, [0, 0, 0, 0])
    refs[5] = f.n_tris + 7  # out of range
    with pytest.raises(rtb200.RtbError) as e:
        m.upload(s.flat)
    assert "reference out of range" in str(e.value)
    refs[5] = keep
    ms = m.upload(s.flat)
    out, _ = ms.render(s.camera, s.setting, rtb200.make_frame(160, 64))
    dev = ctx.upload(s.flat)
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(160, 64))
    assert np.array_equal(_bits(out), _bits(whole))
    dev.close(); ms.close(); m.close(); s.close()


def test_multi_render_all_gpus(ctx):
    n = _n_gpus()
    if n < 2:
        pytest.skip("one GPU on this box (the n-contexts-on-one-device variant covers the code path)")
    w, h = 1280, 960
    s = PresetScene(5, "sah", 150)
    dev = ctx.upload(s.flat)
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    m = rtb200.MultiContext(n)
    ms = m.upload(s.flat)
    pinned = rtb200.PinnedArray((h, w, 3))
    for _ in range(3):
        _, mst = ms.render(s.camera, s.setting, rtb200.make_frame(w, h), out=pinned.array)
        assert np.array_equal(_bits(pinned.array), _bits(whole)) and mst["n_rays"] == st["n_rays"]
    pinned.close(); ms.close(); m.close(); dev.close(); s.close()


def test_multi_render_monte_carlo(ctx):
    w, h = 96, 72
    s = PresetScene(2)
    dev = ctx.upload(s.flat)
    whole, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=4, seed=9))
    m = rtb200.MultiContext(3, [0, 0, 0])
    ms = m.upload(s.flat)
    out, _ = ms.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=4, seed=9))
    assert np.array_equal(_bits(out), _bits(whole))
    ms.close(); m.close(); dev.close(); s.close()


# ---- rtb_ipc_*: the owner's device frame mapped into another PROCESS ----------------------------------------------
_PEER = r"""
import sys, numpy as np
sys.path.insert(0, {root!r})
import rtb200
handle = bytes.fromhex(sys.argv[1]); rank, world, w, h, cb = (int(v) for v in sys.argv[2:7])
ctx = rtb200.Context(0)
s = rtb200.PresetScene(5, "sah", 24)
dev = ctx.upload(s.flat)
frame = ctx.ipc_open(handle)
fr = rtb200.make_frame(w, h, rank=rank, world=world, row_block=8, col_block=cb, layout=rtb200.LAYOUT_GLOBAL)
st = dev.render_device(s.camera, s.setting, fr, frame, 0, want_stats=True)   # stats: synchronises
ctx.ipc_close(frame)
print("PEER_RAYS", st["n_rays"])
"""


@pytest.mark.parametrize("col_block", [0, 8])
def test_ipc_peer_process_stores_into_owner_frame(ctx, col_block):
    """Process A owns the frame (rtb_device_alloc + rtb_ipc_export); process B maps it (rtb_ipc_open) and renders its
    shard straight into it while A renders its own: the frame A reads back is the whole frame.  On a multi-GPU box the
    peers sit on other devices and the stores cross NVLink; the mapping code is the same."""
    w, h, world = 256, 96, 2
    s = PresetScene(5, "sah", 24)
    dev = ctx.upload(s.flat)
    whole, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h))
    nbytes = w * h * 3 * 4
    frame = ctx.device_alloc(nbytes)
    handle = ctx.ipc_export(frame)
    peer = subprocess.Popen([sys.executable, "-c", _PEER.format(root=ROOT), handle.hex(), "1", str(world), str(w), str(h), str(col_block)],
                            stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    fr = rtb200.make_frame(w, h, rank=0, world=world, row_block=8, col_block=col_block, layout=rtb200.LAYOUT_GLOBAL)
    mine = dev.render_device(s.camera, s.setting, fr, frame, 0, want_stats=True)
    out, err = peer.communicate(timeout=300)
    assert peer.returncode == 0, err[-2000:]
    peer_rays = int(out.split("PEER_RAYS")[1].split()[0])
    got = np.zeros((h, w, 3), np.float32)
    ctx.device_download(frame, got)
    assert np.array_equal(_bits(got), _bits(whole))
    assert mine["n_rays"] + peer_rays == st["n_rays"]
    ctx.device_free(frame)
    dev.close(); s.close()


# ---- Monte-Carlo accumulation buffer and sample shards --------------------------------------------------------------
def test_moments_buffer_and_sample_shards(ctx):
    """RTB_OUTPUT_MOMENTS returns (sum, sum of squares) per pixel and channel: sum / spp is the rendered mean (the
    render accumulates radiance * (1 / spp) sample by sample, so equal up to float rounding), the variance estimate is
    non-negative, and the buffers of disjoint SAMPLE shards add up to the buffer of the whole sample range exactly as
    sums of the same per-sample values in another order (float rounding only).  The same for the images of sample
    shards (each sample weighted 1 / spp)."""
    w, h, spp = 64, 48, 24
    s = PresetScene(2)
    dev = ctx.upload(s.flat)
    img, st = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=spp, seed=2))
    mom, mst = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=spp, seed=2, layout=rtb200.OUTPUT_MOMENTS))
    assert mom.shape == (h, w, 6) and mst["n_rays"] == st["n_rays"]
    mean = mom[..., :3] / spp
    assert np.allclose(mean, img, rtol=2e-5, atol=1e-6)
    var = (mom[..., 3:] - mom[..., :3] ** 2 / spp) / (spp - 1)
    assert var.min() > -1e-3 * max(1.0, float(var.max())) and var.mean() > 0
    parts, imgs, rays = [], [], 0
    for first, count in ((0, 8), (8, 8), (16, 8)):
        fm = rtb200.make_frame(w, h, samples=spp, seed=2, layout=rtb200.OUTPUT_MOMENTS, sample_first=first, sample_count=count)
        p, pst = dev.render(s.camera, s.setting, fm)
        parts.append(p.astype(np.float64))
        rays += pst["n_rays"]
        q, _ = dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=spp, seed=2, sample_first=first, sample_count=count))
        imgs.append(q.astype(np.float64))
    assert rays == st["n_rays"]
    assert np.allclose(sum(parts), mom, rtol=2e-6, atol=1e-6)
    assert np.allclose(sum(imgs), img, rtol=2e-5, atol=1e-6)
    with pytest.raises(rtb200.RtbError):  # outside [0, samples)
        dev.render(s.camera, s.setting, rtb200.make_frame(w, h, samples=spp, sample_first=20, sample_count=8))
    with pytest.raises(rtb200.RtbError):  # Whitted has no samples to accumulate
        t = PresetScene(5, "sah", 8)
        d2 = ctx.upload(t.flat)
        try:
            d2.render(t.camera, t.setting, rtb200.make_frame(w, h, layout=rtb200.OUTPUT_MOMENTS))
        finally:
            d2.close(); t.close()
    dev.close(); s.close()


# ---- the reference's callbacks behind the RenderProc -----------------------------------------------------------------
@pytest.mark.parametrize("n_devices", [1, 2])
def test_progress_and_log_callbacks(ctx, n_devices):
    """Script::Run(CudaRenderer::Render, alg, log, progress, ...): the ProgressCallback is called in the reference's unit
    (rows of `height`, MainWindow.cpp:271), monotonically, ends at (height, height); a failing render reports through
    the LogCallback and returns a negative time (Scripts.h:9-12 has no other error channel)."""
    if n_devices > _n_gpus():
        pytest.skip("needs more GPUs")
    w, h = 1600, 1200
    img, info = rtb200.script_run_ex(5, "sah", 150, w, h, n_devices=n_devices)
    assert info["rc"] == 0 and img is not None
    assert info["progress_calls"] >= 1 and info["progress_monotone"]
    assert (info["progress_last"], info["progress_total"]) == (h, h)
    ref, _ = rtb200.script_run(5, "sah", 150, w, h)
    assert np.array_equal(_bits(img), _bits(ref))
    # failure path: a frame size the library refuses
    img, info = rtb200.script_run_ex(5, "sah", 12, 0, 10, n_devices=n_devices)
    assert info["rc"] < 0 and img is None and "CudaRenderer" in info["log"]


def test_first_frame_schedule_can_be_forgotten(ctx):
    """rtb_forget_schedule: the next frame runs in raster order through one throughput kernel like the first frame of a
    view (fewer launches than a tiered frame) and produces the same image."""
    w, h = 1280, 960
    s = PresetScene(5, "sah", 150)
    dev = ctx.upload(s.flat)
    fr = rtb200.make_frame(w, h)
    ctx.forget_schedule()
    first, st1 = dev.render(s.camera, s.setting, fr)
    second, st2 = dev.render(s.camera, s.setting, fr)
    third, st3 = dev.render(s.camera, s.setting, fr)
    assert st1["n_launches"] == 4 and st3["n_launches"] > 4  # 1 render + 3 order kernels; then the tiers join
    ctx.forget_schedule()
    again, st4 = dev.render(s.camera, s.setting, fr)
    assert st4["n_launches"] == 4
    for img in (second, third, again):
        assert np.array_equal(_bits(img), _bits(first))
    dev.close(); s.close()


def test_frames_on_two_caller_streams_of_one_context(ctx):
    """The tile schedule (cost / order buffers, side stream) belongs to the context and is shared by its frames; a frame
    issued on another caller stream is ordered behind the previous one by the library (an event recorded behind each frame's
    last kernel), including when the tile buffers have to grow.  Alternating streams and frame sizes must give, every
    time, the image a fresh context renders."""
    import torch
    s = PresetScene(5, "sah", 40)
    dev = ctx.upload(s.flat)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    sizes = [(256, 192), (512, 384), (256, 192), (640, 480), (512, 384), (640, 480)]
    want = {}
    for (w, h) in set(sizes):
        c2 = rtb200.Context(0)
        d2 = c2.upload(s.flat)
        want[(w, h)], _ = d2.render(s.camera, s.setting, rtb200.make_frame(w, h))
        d2.close(); c2.close()
    bufs = []
    for i, (w, h) in enumerate(sizes * 2):
        st = streams[i % 2]
        with torch.cuda.stream(st):
            buf = torch.empty((h, w, 3), dtype=torch.float32, device="cuda:0")
        st.synchronize()
        dev.render_device(s.camera, s.setting, rtb200.make_frame(w, h), buf.data_ptr(), st.cuda_stream)  # no sync in between
        bufs.append((w, h, buf))
    torch.cuda.synchronize()
    for (w, h, buf) in bufs:
        assert np.array_equal(_bits(buf.cpu().numpy()), _bits(want[(w, h)]))
    dev.close(); s.close()

