"""N > 1 host logic on CPU (gloo, world_size 2 and 3): the tile-row shard arithmetic of include/rtb.h
(rtb_shard_rows, block-cyclic dealing), the padded all-gather layout bench.py uses, and the row order
rtb_unshard_device restores; the same for column-block shards (rtb_frame.col_block, rtb_shard_width,
rtb_unshard_cols_device).  No GPU kernels run here; each rank fills its shard with a function of the
global row index so the reassembled frame can be checked exactly."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import numpy as np, torch, torch.distributed as dist
    import rtb200
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    W, H, RB = 40, 300, 16
    frame = rtb200.make_frame(W, H, rank=rank, world=world, row_block=RB)
    rows = rtb200.shard_rows(frame)                         # C ABI: rtb_shard_rows
    ys = rtb200.shard_row_indices(H, rank, world, RB)
    assert rows == len(ys)
    t = torch.tensor([rows], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rows_max = int(t.item())
    local = torch.zeros((rows_max, W, 3), dtype=torch.float32)
    xs = torch.arange(W, dtype=torch.float32)
    for lr, y in enumerate(ys):                              # "render": pixel value encodes (y, x, channel)
        for c in range(3):
            local[lr, :, c] = float(y) * 1000 + xs + c * 0.25
    gathered = torch.zeros((world, rows_max, W, 3), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered.view(world * rows_max, W, 3), local)
    # the permutation rtb_unshard_device applies (k_unshard): row y lives at [b %% world][ (b // world)*RB + y - b*RB ]
    image = torch.empty((H, W, 3), dtype=torch.float32)
    for y in range(H):
        b = y // RB
        image[y] = gathered[b %% world, (b // world) * RB + (y - b * RB)]
    expect = torch.arange(H, dtype=torch.float32)[:, None, None] * 1000 + xs[None, :, None] + torch.tensor([0, 0.25, 0.5])[None, None, :]
    assert torch.equal(image, expect), "reassembled frame differs"
    total = torch.tensor([rows], dtype=torch.int64)
    dist.all_reduce(total)
    assert int(total.item()) == H

    # ---- column-block shards (rtb_frame.col_block): local image [H][W / world], gathered [world][H][W / world][3] ----
    CB = 8
    Wc = world * CB * 5
    cframe = rtb200.make_frame(Wc, H, rank=rank, world=world, row_block=RB, col_block=CB)
    assert rtb200.shard_rows(cframe) == H and rtb200.shard_width(cframe) == Wc // world    # C ABI
    Wl = Wc // world
    local = torch.zeros((H, Wl, 3), dtype=torch.float32)
    for y in range(H):
        gx = torch.from_numpy(rtb200.shard_col_indices(Wc, y, rank, world, RB, CB)).to(torch.float32)
        for c in range(3):
            local[y, :, c] = float(y) * 1000 + gx + c * 0.25
    gathered = torch.zeros((world, H, Wl, 3), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered.view(world * H, Wl, 3), local)
    # the permutation rtb_unshard_cols_device applies (k_unshard_cols): column block bx of row y lives on rank
    # (bx + y // RB) %% world at local block bx // world
    image = torch.empty((H, Wc, 3), dtype=torch.float32)
    for y in range(H):
        for bx in range(Wc // CB):
            r = (bx + y // RB) %% world
            image[y, bx * CB:(bx + 1) * CB] = gathered[r, y, (bx // world) * CB:(bx // world + 1) * CB]
    xc = torch.arange(Wc, dtype=torch.float32)
    expect = torch.arange(H, dtype=torch.float32)[:, None, None] * 1000 + xc[None, :, None] + torch.tensor([0, 0.25, 0.5])[None, None, :]
    assert torch.equal(image, expect), "column-sharded frame differs"

    # ---- RTB_LAYOUT_GLOBAL: every rank stores its own pixels into ONE whole frame; no gather, no unshard.  Here each
    # rank fills a private copy of the frame at the slots it owns (shard_pixel_mask = the kernels' localToGlobal) and a
    # reduce(sum) onto rank 0 stands in for the shared buffer: the supports must be disjoint and cover the frame ----
    for (cb, Wg) in ((0, W), (CB, Wc)):
        mask = torch.from_numpy(rtb200.shard_pixel_mask(Wg, H, rank, world, RB, cb))
        assert int(mask.sum()) == rtb200.shard_rows(rtb200.make_frame(Wg, H, rank=rank, world=world, row_block=RB, col_block=cb)) * \
            rtb200.shard_width(rtb200.make_frame(Wg, H, rank=rank, world=world, row_block=RB, col_block=cb))
        xg = torch.arange(Wg, dtype=torch.float32)
        full = torch.arange(H, dtype=torch.float32)[:, None, None] * 1000 + xg[None, :, None] + torch.tensor([0, 0.25, 0.5])[None, None, :]
        mine = torch.where(mask[:, :, None], full, torch.zeros_like(full))
        owners = mask.to(torch.int32)
        dist.reduce(mine, dst=0, op=dist.ReduceOp.SUM)
        dist.reduce(owners, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            assert bool((owners == 1).all()), "shards overlap or leave holes"
            assert torch.equal(mine, full), "global-layout frame differs"

    # ---- Monte-Carlo SAMPLE shards (rtb_frame.sample_first / sample_count): the ranges partition [0, samples), and the
    # per-rank images (each sample weighted 1 / samples) sum to the frame under reduce(sum) ----
    for samples in (64, 7, world):
        first, count = rtb200.sample_shard(samples, rank, world)
        cover = torch.zeros(samples, dtype=torch.int32)
        cover[first:first + count] = 1
        dist.all_reduce(cover)
        assert bool((cover == 1).all()) and count >= samples // world
        vals = torch.arange(samples, dtype=torch.float64) + 1.0            # "radiance" of sample s at one pixel
        part = (vals[first:first + count] / samples).sum().reshape(1)
        dist.reduce(part, dst=0, op=dist.ReduceOp.SUM)
        if rank == 0:
            assert abs(float(part) - float(vals.mean())) < 1e-12
    dist.barrier()
    if rank == 0:
        print("GLOO_OK", world, rows_max)
    dist.destroy_process_group()
""")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_frame_reassembles_over_gloo(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = _free_port()
    procs = []
    for rank in range(world):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                   MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert any("GLOO_OK" in o for o in outs)
