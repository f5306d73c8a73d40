mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "roi_align" > gpurun_out/pytest_fwd.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_fwd.log
timeout 300 python tools/micro_roi.py > gpurun_out/micro16.log 2>&1; tail -1 gpurun_out/micro16.log
timeout 300 python tools/micro_roi.py --bf16 > gpurun_out/micro16_bf16.log 2>&1; tail -1 gpurun_out/micro16_bf16.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_band -c 1 -o gpurun_out/roi_fwd_v7 -f python tools/roi_only.py fwd > gpurun_out/ncu_fwd7.log 2>&1; tail -1 gpurun_out/ncu_fwd7.log
