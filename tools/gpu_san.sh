mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "roi_align and not full_size" > gpurun_out/sanitizer_roi.log 2>&1; echo "memcheck rc=$?"
tail -8 gpurun_out/sanitizer_roi.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py tests/test_gemm_gpu.py -m gpu -q -x -k "not roi_align" > gpurun_out/sanitizer_rest.log 2>&1; echo "memcheck(rest) rc=$?"
tail -6 gpurun_out/sanitizer_rest.log
