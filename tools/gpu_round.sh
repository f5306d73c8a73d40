mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest9.log 2>&1; echo "pytest rc=$?"
timeout 200 python tools/gemm_diag.py > gpurun_out/gemm_diag2.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench4.log 2>&1
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench4_ref.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_list2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_fwd_slab2 -c 1 -o gpurun_out/roi_fwd_v3 -f python tools/roi_only.py fwd > gpurun_out/ncu_fwd3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_slab2 -c 1 -o gpurun_out/roi_bwd_v4 -f python tools/roi_only.py bwd > gpurun_out/ncu_bwd4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tf32_gemm_kernel -c 1 -o gpurun_out/gemm_tf32 -f python tools/gemm_only.py > gpurun_out/ncu_gemm.log 2>&1
tail -3 gpurun_out/pytest9.log; cat gpurun_out/gemm_diag2.log | tail -20; tail -2 gpurun_out/bench4.log; tail -1 gpurun_out/bench4_ref.log
