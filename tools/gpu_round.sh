#!/bin/bash
# One GPU-box pass of everything the round is judged on (run through gpurun; writes into gpurun_out/):
#   parity tests, both bench arms, the ncu launch list of the bench, ncu --set full of the two ROIAlign kernels and the
#   tcgen05 GEMM, compute-sanitizer over the op tests.  Summaries for profiles/ come from tools/ncu_summary.py.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_final.log
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_final.log 2>&1; tail -1 gpurun_out/bench_final.log | cut -c1-400
timeout 400 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/bench_final_ref.log 2>&1; tail -1 gpurun_out/bench_final_ref.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final.log 2>&1; tail -1 gpurun_out/smoke_final.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/ncu_list_final.log 2>&1
if [ -z "$SKIP_NCU_FULL" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_band -c 1 -o gpurun_out/roi_fwd_final -f python tools/roi_only.py fwd > gpurun_out/ncu_fwd_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_cl -c 1 -o gpurun_out/roi_bwd_final -f python tools/roi_only.py bwd > gpurun_out/ncu_bwd_final.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tf32_gemm_kernel -c 1 -o gpurun_out/gemm_final -f python tools/gemm_only.py > gpurun_out/ncu_gemm_final.log 2>&1
fi
timeout 120 python tools/micro_roi.py > gpurun_out/micro_roi_f32.log 2>&1; tail -1 gpurun_out/micro_roi_f32.log | cut -c1-160
timeout 120 python tools/micro_roi.py --bf16 > gpurun_out/micro_roi_bf16.log 2>&1; tail -1 gpurun_out/micro_roi_bf16.log
timeout 120 python tools/micro_weak.py 2 2000 20 > gpurun_out/micro_weak.log 2>&1; tail -1 gpurun_out/micro_weak.log
timeout 400 python bench.py --steps 30 --warmup 5 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_final_bf16.log 2>&1; tail -1 gpurun_out/bench_final_bf16.log | cut -c1-200
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py tests/test_weak_gpu.py -m gpu -q -k "not full_size and not full_batch" > gpurun_out/sanitizer_final.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_final.log
