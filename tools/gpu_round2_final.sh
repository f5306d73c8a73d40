#!/bin/bash
# Round-2 final evidence pass (1 GPU): full GPU suite, every bench arm, smoke, launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final3.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final3.log 2>&1; tail -1 gpurun_out/r2_bench_final3.log | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline --no-aux > gpurun_out/r2_bench_final3_bf16.log 2>&1; tail -1 gpurun_out/r2_bench_final3_bf16.log | cut -c1-250
timeout 300 python bench.py --mode infer --steps 50 --warmup 5 > gpurun_out/r2_bench_final3_infer.log 2>&1; tail -1 gpurun_out/r2_bench_final3_infer.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final3.log 2>&1; tail -1 gpurun_out/r2_smoke_final3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/r2_launches_final3.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2_ncu_launches_final3.log 2>&1; echo "launch list rc=$?"
