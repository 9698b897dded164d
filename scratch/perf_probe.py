import sys, os
sys.path.insert(0, '.')
import torch, numpy as np, rtb200
ctx = rtb200.Context(0)
W, H = 3840, 2880
buf = torch.empty((H, W, 3), dtype=torch.float32, device="cuda:0")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
tag = os.environ.get("TAG", "")
for preset, alg in [(5, "sah"), (5, "rgrid"), (4, "sah")]:
    s = rtb200.PresetScene(preset, alg, 150)
    d = ctx.upload(s.flat)
    out = []
    for world in (1, 8):
        per = []
        for rank in range(world):
            fr = rtb200.make_frame(W, H, rank=rank, world=world, row_block=16)
            times = [d.render_device(s.camera, s.setting, fr, buf.data_ptr(), st.cuda_stream, want_stats=True)["kernel_ms"] for _ in range(4)]
            per.append((times[0], min(times[1:])))
        out.append((world, [round(a, 2) for a, b in per], [round(b, 2) for a, b in per]))
    print(tag, preset, alg, "N=1 first %.2f ordered %.2f | N=8 shards first max %.2f ordered max %.2f (all %s)" % (out[0][1][0], out[0][2][0], max(out[1][1]), max(out[1][2]), out[1][2]))
    d.close(); s.close()
