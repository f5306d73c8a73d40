#!/bin/bash
# Round-2 evidence pass C: full GPU suite, every bench arm, smoke, ncu of the small kernels and of the GEMMs.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_final.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_final.log 2>&1; tail -1 gpurun_out/r2_bench_final.log | cut -c1-330
timeout 300 python bench.py --steps 20 --warmup 5 --dtype bf16 --no-cpu-baseline --no-aux > gpurun_out/r2_bench_final_bf16.log 2>&1; tail -1 gpurun_out/r2_bench_final_bf16.log | cut -c1-250
timeout 300 python bench.py --mode infer --steps 20 --warmup 5 > gpurun_out/r2_bench_final_infer.log 2>&1; tail -1 gpurun_out/r2_bench_final_infer.log | cut -c1-250
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_final.log 2>&1; tail -1 gpurun_out/r2_smoke_final.log
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,launch__grid_size,launch__block_size,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum"
timeout 300 ncu --metrics $M --clock-control none -k regex:"nms_|filter_|softmax_decode|segmented" -s 12 -c 36 --csv --log-file gpurun_out/r2_small_detect.csv python tools/micro_detect.py > gpurun_out/r2_micro_detect.log 2>&1; echo "detect ncu rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k regex:"mask_" -c 12 --csv --log-file gpurun_out/r2_small_mask.csv python tools/micro_mask.py > gpurun_out/r2_micro_mask.log 2>&1; echo "mask ncu rc=$?"
timeout 300 ncu --metrics $M --clock-control none -k regex:"iou_match|label_kernel|sample_gather" -s 3 -c 6 --csv --log-file gpurun_out/r2_small_match.csv python tools/match_only.py 64 > gpurun_out/r2_match_only.log 2>&1; echo "match ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"tf32_gemm_kernel|tf32_wgrad_kernel" -s 2 -c 2 -o gpurun_out/r2_gemm -f python tools/gemm_only.py > gpurun_out/r2_ncu_gemm.log 2>&1; echo "gemm ncu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_band -c 1 -o gpurun_out/r2_fwd_final -f python tools/roi_only.py fwd > gpurun_out/r2_ncu_fwd_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_cl2 -c 1 -o gpurun_out/r2_bwd_final -f python tools/roi_only.py bwd > gpurun_out/r2_ncu_bwd_final.log 2>&1
python tools/micro_mask.py 2>&1 | tail -1
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_final_ref.log 2>&1; tail -1 gpurun_out/r2_bench_final_ref.log | cut -c1-900
