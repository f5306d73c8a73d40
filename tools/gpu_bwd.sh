mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_own -c 1 -o gpurun_out/roi_bwd_own1 -f python tools/roi_only.py bwd > gpurun_out/ncu_bwd_own1.log 2>&1; tail -1 gpurun_out/ncu_bwd_own1.log
