#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_bench_config_gpu.py -m gpu -q -k "roi or bench_config or full_size" 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r2_bench_h.log 2>&1; tail -1 gpurun_out/r2_bench_h.log | cut -c1-260
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"roi_tables_kernel|bwd_tables" -c 6 --csv --log-file gpurun_out/r2_tables_h.csv python tools/roi_only.py fwd > /dev/null 2>&1; grep -o '"gpu__time_duration.sum","[a-z]*","[0-9.]*"' gpurun_out/r2_tables_h.csv | head -4
