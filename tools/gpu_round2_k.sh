#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gemm_gpu.py -m gpu -q 2>&1 | tail -1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:wgrad_reduce -s 2 -c 1 -o gpurun_out/r2_wgrad_reduce -f python tools/gemm_only.py > /dev/null 2>&1; echo "ncu rc=$?"
