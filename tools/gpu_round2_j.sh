#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gemm_gpu.py tests/test_bench_config_gpu.py tests/test_heads_gpu.py -m gpu -q -x > gpurun_out/r2_pytest_j.log 2>&1; tail -40 gpurun_out/r2_pytest_j.log | cut -c1-220
