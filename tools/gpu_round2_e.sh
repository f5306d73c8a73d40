#!/bin/bash
# Round-2 pass E: unrolled backward sweeps + separable mask paste: parity tests, micro timings.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_bench_config_gpu.py -m gpu -q -k "roi or paste or backward or bwd or full_size" > gpurun_out/r2_pytest_e.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_e.log
echo "== micro_roi f32"; timeout 200 python tools/micro_roi.py 2>&1 | tail -1
echo "== micro_mask"; timeout 200 python tools/micro_mask.py 2>&1 | tail -1
