#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gemm_gpu.py tests/test_bench_config_gpu.py tests/test_heads_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-aux > gpurun_out/r2_bench_i.log 2>&1; tail -1 gpurun_out/r2_bench_i.log | cut -c1-260
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tf32_gemm|splitk_reduce|tf32_wgrad|wgrad_reduce|bwd_prepare|unpermute|roi_tables|similarity_transfer|fastrcnn_loss|loss_reduce" -c 60 --csv --log-file gpurun_out/r2_small_i.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r2_small_i.csv')) if len(r)>5]
h=rows[0]; acc=collections.defaultdict(list)
for r in rows[1:]:
    d=dict(zip(h,r))
    if d['Metric Name']=='gpu__time_duration.sum': acc[d['Kernel Name'].split('(')[0][-40:]].append(float(d['Metric Value'])/1000)
for k,v in acc.items(): print(k, len(v), round(sum(v)/len(v),2))
PY
