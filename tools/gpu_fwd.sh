mkdir -p gpurun_out
UNIT_ROI_FWD_CL=1 timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "roi_align" > gpurun_out/pytest_fwdcl.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_fwdcl.log
UNIT_ROI_FWD_CL=1 timeout 120 python tools/micro_roi.py > gpurun_out/micro20.log 2>&1; tail -1 gpurun_out/micro20.log
UNIT_ROI_FWD_CL=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_cl -c 1 -o gpurun_out/roi_fwd_cl2 -f python tools/roi_only.py fwd > gpurun_out/ncu_fwd_cl2.log 2>&1; tail -1 gpurun_out/ncu_fwd_cl2.log
