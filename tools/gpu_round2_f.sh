#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest_g.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_g.log
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r2_bench_g.log 2>&1; tail -1 gpurun_out/r2_bench_g.log | cut -c1-330
timeout 300 python bench.py --mode infer --steps 50 --warmup 5 > gpurun_out/r2_bench_infer_g.log 2>&1; grep '^{' gpurun_out/r2_bench_infer_g.log | python -c "import sys,json;d=json.loads(sys.stdin.read());print('infer',round(d['value']),round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],4))"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_g.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2_ncu_launches_g.log 2>&1; echo "launch list rc=$?"
