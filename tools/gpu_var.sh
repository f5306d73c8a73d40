mkdir -p gpurun_out
for v in nored sametile both; do
  echo "== $v"; UNIT_B200_LIB=$PWD/unit_b200/build/variants/lib_$v.so timeout 200 python tools/micro_roi.py 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ours_bwd_ms'], d['ours_fwd_ms'])"
done
