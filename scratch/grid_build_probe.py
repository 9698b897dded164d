import sys, time
sys.path.insert(0, '.')
import rtb200
ctx = rtb200.Context(0)
for preset, alg, seg in [(5, "rgrid", 150), (5, "fgrid", 150), (4, "rgrid", 150), (4, "fgrid", 150)]:
    t0 = time.perf_counter(); host = rtb200.PresetScene(preset, alg, seg); t1 = time.perf_counter()
    rtb200.set_grid_on_device(True)
    lazy = rtb200.PresetScene(preset, alg, seg); t2 = time.perf_counter()
    rtb200.set_grid_on_device(False)
    ups = []
    for s in (host, lazy):
        best = 1e9
        for _ in range(4):
            a = time.perf_counter(); d = ctx.upload(s.flat); b = time.perf_counter(); best = min(best, (b - a) * 1e3)
            h = d.grid_hash(); d.close()
        ups.append((best, h))
    print(f"p{preset} {alg}: scene + HOST grid build {1e3*(t1-t0):.0f} ms (host prepare {host.prepare_ms} ms), scene without grid {1e3*(t2-t1):.0f} ms | upload of host-built {ups[0][0]:.2f} ms, upload + DEVICE build {ups[1][0]:.2f} ms | identical {ups[0][1] == ups[1][1]} {ups[1][1][1]}", flush=True)
    host.close(); lazy.close()
