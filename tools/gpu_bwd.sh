mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -k "roi_align" > gpurun_out/pytest_bwd.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_bwd.log
timeout 300 python tools/micro_roi.py > gpurun_out/micro17.log 2>&1; tail -1 gpurun_out/micro17.log
UNIT_ROI_BWD_BAND=0 timeout 300 python tools/micro_roi.py > gpurun_out/micro17_noband.log 2>&1; tail -1 gpurun_out/micro17_noband.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_cl -c 1 -o gpurun_out/roi_bwd_v8 -f python tools/roi_only.py bwd > gpurun_out/ncu_bwd8.log 2>&1; tail -1 gpurun_out/ncu_bwd8.log
