#!/bin/bash
# Round-2 pass D: 13-warp backward variant, bf16 ROIAlign re-check, mask paste timing, backward ncu with the final L2 policy,
# inference arm with the end-of-loop gather, full GPU suite, default bench.
mkdir -p gpurun_out
echo "== micro_roi f32 (default lib)"; timeout 200 python tools/micro_roi.py 2>&1 | tail -1
echo "== micro_roi f32 (13 warps)"; UNIT_B200_LIB=$PWD/unit_b200/build/variants/lib_nw13.so timeout 200 python tools/micro_roi.py 2>&1 | tail -1
echo "== micro_roi bf16"; timeout 200 python tools/micro_roi.py --bf16 2>&1 | tail -1
echo "== micro_mask"; timeout 200 python tools/micro_mask.py 2>&1 | tail -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:roi_align_bwd_cl2 -c 1 -o gpurun_out/r2_bwd_final2 -f python tools/roi_only.py bwd > gpurun_out/r2_ncu_bwd_final2.log 2>&1; echo "bwd ncu rc=$?"
timeout 300 python bench.py --mode infer --steps 20 --warmup 5 > gpurun_out/r2_bench_infer2_n1.log 2>&1; tail -1 gpurun_out/r2_bench_infer2_n1.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_d.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_d.log 2>&1; tail -1 gpurun_out/r2_bench_d.log | cut -c1-400
