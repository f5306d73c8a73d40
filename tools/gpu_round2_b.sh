#!/bin/bash
# Round-2 evidence pass B: sanitizer (racecheck / initcheck / memcheck) over the op tests, mask micro, both bench arms.
mkdir -p gpurun_out
python tools/micro_mask.py 2>&1 | tail -1
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "mask" 2>&1 | tail -2
SEL="roi_align_backward_channel_lane or roi_align_backward_sorted or roi_align_forward_slab or nms or detect or matcher or label or softmax or transfer or weak or wgrad or gemm2"
for tool in racecheck initcheck memcheck; do
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 7 python -m pytest tests/test_ops_gpu.py tests/test_weak_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "($SEL) and not full_size and not full_batch and not opcheck" > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?"; tail -3 gpurun_out/r2_sanitizer_$tool.log | cut -c1-200
done
