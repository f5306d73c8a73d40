#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "paste" 2>&1 | tail -1
timeout 200 python tools/paste_probe.py 2>&1 | tail -1
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__registers_per_thread,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"
timeout 300 ncu --metrics $M --clock-control none -k regex:mask_paste_rows -c 1 --csv --log-file gpurun_out/r2_paste_probe.csv python tools/paste_probe.py --once > /dev/null 2>&1; echo "ncu rc=$?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_paste_probe.csv')) if len(r)>5]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['ID'], d['Metric Name'], d['Metric Value'], d['Metric Unit'])
PY
