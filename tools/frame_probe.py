"""Kernel milliseconds of bench workloads on one GPU: whole frames, or the N shards of a frame one after the other
(what bounds the N-GPU step: the slowest rank's kernels).

  python tools/frame_probe.py --workloads p5_sah_4k,p5_rgrid_4k [--world 8 --col-block 32] [--sizes 400x300,3840x2880] [--reps 8]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rtb200  # noqa: E402
from bench import WORKLOADS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workloads", default="p5_sah_4k")
ap.add_argument("--sizes", default="")
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--col-block", type=int, default=0)
ap.add_argument("--reps", type=int, default=8)
ap.add_argument("--tag", default=os.environ.get("TAG", ""))
ap.add_argument("--global-frame", action="store_true", help="shards store into ONE whole device frame (RTB_LAYOUT_GLOBAL), as the ranks of bench.py do")
args = ap.parse_args()
ctx = rtb200.Context(0)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for name in args.workloads.split(","):
    wl = WORKLOADS[name]
    s = rtb200.PresetScene(wl["preset"], wl["algorithm"], wl["segments"])
    setting = rtb200.make_setting(wl["setting"]) if "setting" in wl else s.setting
    d = ctx.upload(s.flat)
    sizes = [tuple(int(v) for v in z.split("x")) for z in args.sizes.split(",")] if args.sizes else [(wl["width"], wl["height"])]
    for (W, H) in sizes:
        per_rank = []
        for rank in range(args.world):
            fr = rtb200.make_frame(W, H, samples=wl["samples"], rank=rank, world=args.world, col_block=args.col_block if args.world > 1 else 0,
                                   layout=rtb200.LAYOUT_GLOBAL if args.global_frame else 0)
            buf = torch.empty((H, W, 3) if args.global_frame else (max(rtb200.shard_rows(fr), 1), rtb200.shard_width(fr), 3), dtype=torch.float32, device="cuda:0")
            ctx.forget_schedule()
            t = [d.render_device(s.camera, setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True) for _ in range(args.reps)]
            ms = [x["kernel_ms"] for x in t]
            per_rank.append((ms[0], float(np.min(ms[3:])) if len(ms) > 3 else ms[-1], t[-1]["n_rays"]))
        rays = sum(p[2] for p in per_rank)
        worst = max(p[1] for p in per_rank)
        print(args.tag, name, f"{W}x{H} world {args.world} col_block {args.col_block}: first-frame ms", " ".join(f"{p[0]:.2f}" for p in per_rank),
              "| steady ms", " ".join(f"{p[1]:.2f}" for p in per_rank), f"| slowest {worst:.2f} ms -> {rays / worst / 1e3:.0f} Mrays/s if the ranks ran side by side", flush=True)
    d.close(); s.close()
ctx.close()
