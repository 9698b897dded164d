timeout 600 python -m pytest tests/test_gpu_frame.py tests/test_route_a.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/layout_probe.py 2>&1 | tee gpurun_out/layout_probe2.log
timeout 200 python tools/layout_probe.py --workload p5_rgrid_4k 2>&1 | tee -a gpurun_out/layout_probe2.log
