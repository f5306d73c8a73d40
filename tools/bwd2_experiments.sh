#!/bin/bash
# Backward (cl2) experiments: read+red overlap micro-benchmark, kernel variants, L2 policies.
mkdir -p gpurun_out
tools/micro/read_red_mix
run() { echo -n "$1: "; env $2 python tools/micro_roi.py 2>/dev/null | tail -1 | python -c "import sys,json;d=json.load(sys.stdin);print('fwd %.4f bwd %.4f'%(d['ours_fwd_ms'],d['ours_bwd_ms']))"; }
V=unit_b200/build/variants
run "shipped" ""
run "shipped evict_normal" "UNIT_ROI_BWD_EVICT_FIRST=0"
run "red evict_last hint" "UNIT_B200_LIB=$V/lib_redlast.so"
run "red evict_last hint + tiles evict_normal" "UNIT_B200_LIB=$V/lib_redlast.so UNIT_ROI_BWD_EVICT_FIRST=0"
run "red hint + normal + promo 64B" "UNIT_B200_LIB=$V/lib_redlast.so UNIT_ROI_BWD_EVICT_FIRST=0 UNIT_ROI_BWD_PROMO=1"
run "no reductions" "UNIT_B200_LIB=$V/lib_nored.so"
run "10 warps" "UNIT_B200_LIB=$V/lib_nw10.so"
run "8 warps" "UNIT_B200_LIB=$V/lib_nw8.so"
export UNIT_B200_LIB=$V/lib_redlast.so UNIT_ROI_BWD_EVICT_FIRST=0
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:roi_align_bwd_cl2 -c 1 --csv python tools/roi_only.py bwd 2>/dev/null | grep -E "dram__bytes|lts__t_bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s ", $(NF-2), $NF, $(NF-1)}' | tr -d '"'; echo " (red hint + evict_normal)"
