#!/bin/bash
# Round-2 evidence pass A: new tests, launch list of the bench step, basic ncu metrics of the inference-side kernels.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_heads_gpu.py -m gpu -q -k "outputs_variants" > gpurun_out/r2_pytest9.log 2>&1; echo "outputs test rc=$?"; tail -3 gpurun_out/r2_pytest9.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/r2_ncu_list.log 2>&1; echo "launch list rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:"nms_|detect_filter|softmax_decode|segmented" -c 60 --csv --log-file gpurun_out/r2_small_detect.csv python tools/micro_detect.py > gpurun_out/r2_micro_detect.log 2>&1; echo "detect ncu rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k regex:"mask_" -c 12 --csv --log-file gpurun_out/r2_small_mask.csv python tools/micro_mask.py > gpurun_out/r2_micro_mask.log 2>&1; echo "mask ncu rc=$?"
python tools/micro_detect.py 2>&1 | tail -1
python tools/micro_mask.py 2>&1 | tail -2
timeout 400 python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline --no-aux > gpurun_out/r2_bench_bf16.log 2>&1; tail -1 gpurun_out/r2_bench_bf16.log | cut -c1-250
