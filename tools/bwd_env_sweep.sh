#!/bin/bash
# Backward channel-lane kernel: time + DRAM bytes under the L2 policy / promotion switches (one ncu pass each).
mkdir -p gpurun_out
for cfg in "1 2" "0 2" "1 0" "0 0" "1 3" "0 3" "0 1"; do
  set -- $cfg
  export UNIT_ROI_BWD_EVICT_FIRST=$1 UNIT_ROI_BWD_PROMO=$2
  t=$(python tools/micro_roi.py 2>/dev/null | tail -1 | python -c "import sys,json;print(json.load(sys.stdin)['ours_bwd_ms'])")
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,gpu__time_duration.sum --clock-control none -k regex:roi_align_bwd_cl -c 1 --csv python tools/roi_only.py bwd 2>/dev/null | grep -E "dram__bytes|lts__t_bytes|gpu__time" | awk -F'","' '{printf "%s=%s%s ", $(NF-2), $NF, $(NF-1)}' | tr -d '"'
  echo " evict_first=$1 promo=$2 op_ms=$t"
done
