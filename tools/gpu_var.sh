mkdir -p gpurun_out
for promo in 0 1 2 3; do for ef in 0 1; do
  echo "== promo $promo evict_first $ef"; UNIT_ROI_BWD_PROMO=$promo UNIT_ROI_BWD_EVICT_FIRST=$ef timeout 200 python tools/micro_roi.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ours_bwd_ms'])"
done; done
