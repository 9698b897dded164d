"""The command profiled under ncu: warm frames of ONE bench workload, device-resident, then nothing else.

  python tools/profile_frame.py --workload p5_sah_4k [--frames 5]

Frame 1 of a view runs in raster order (one render kernel); from frame 2 on the tile order and the latency tiers are in
play (k-d / grid frames: up to three render kernels side by side).  `profiles/r02_collect.py` takes the render kernels of
the LAST frame from the ncu report, i.e. a steady-state frame like the ones bench.py times.  Prints the kernel time and
the launch count of every frame so the report can be matched to them.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import rtb200  # noqa: E402
from bench import WORKLOADS  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="p5_sah_4k")
ap.add_argument("--frames", type=int, default=5)
ap.add_argument("--rank", type=int, default=0)
ap.add_argument("--world", type=int, default=1)
ap.add_argument("--col-block", type=int, default=0)
ap.add_argument("--host-frame", action="store_true", help="rtb_render into a page-locked host frame (the kernels store over PCIe, in blocks)")
args = ap.parse_args()
wl = WORKLOADS[args.workload]
ctx = rtb200.Context(0)
s = rtb200.PresetScene(wl["preset"], wl["algorithm"], wl["segments"])
setting = rtb200.make_setting(wl["setting"]) if "setting" in wl else s.setting
d = ctx.upload(s.flat)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
fr = rtb200.make_frame(wl["width"], wl["height"], samples=wl["samples"], rank=args.rank, world=args.world, col_block=args.col_block)
buf = torch.empty((rtb200.shard_rows(fr), rtb200.shard_width(fr), 3), dtype=torch.float32, device="cuda:0")
pinned = rtb200.PinnedArray((wl["height"], wl["width"], 3)) if args.host_frame else None
for i in range(args.frames):
    if args.host_frame:
        _, r = d.render(s.camera, setting, fr, out=pinned.array)
        print(f"frame {i}: kernel_ms {r['kernel_ms']:.3f} launches {r['n_launches']} rays {r['n_rays']} (host frame)", flush=True)
        continue
    r = d.render_device(s.camera, setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)
    print(f"frame {i}: kernel_ms {r['kernel_ms']:.3f} launches {r['n_launches']} rays {r['n_rays']}", flush=True)
d.close(); s.close(); ctx.close()
