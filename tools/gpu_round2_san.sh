#!/bin/bash
# compute-sanitizer over the kernels added in the second half of round 2: mask_paste_rows, append_gt, sample_gather with field,
# bwd_prepare + unrolled backward sweeps, GEMM operand ring, warp-parallel lane map (forward tables).
mkdir -p gpurun_out
SEL='(paste and not full_size) or append_gt or gemm2 or predictor_gemm or wgrad or roi_align_backward_channel_lane or roi_align_forward_slab or label'
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -m gpu -q -k "$SEL and not opcheck and not full_batch" > gpurun_out/r2_san2_$tool.log 2>&1
  echo "$tool rc=$?"; grep -E "passed|failed|ERROR SUMMARY|RACECHECK SUMMARY" gpurun_out/r2_san2_$tool.log | tail -3
done
