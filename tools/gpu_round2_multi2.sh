#!/bin/bash
# Round-2 multi-GPU pass 2 (gpurun --gpus 8): inference arm with the end-of-loop gather at 8 / 4 / 2 / 1; paste strip variants.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --mode infer --gpus $n --steps 50 --warmup 5 > gpurun_out/r2_bench_infer2_n$n.log 2>&1
  echo "infer n=$n rc=$?"; grep '^{' gpurun_out/r2_bench_infer2_n$n.log | tail -1 | python -c "import sys,json;d=json.loads(sys.stdin.read());print(d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],4),'dets',d['detections_last_step'],d.get('detections_gathered'))"
done
timeout 300 python bench.py --mode infer --steps 50 --warmup 5 > gpurun_out/r2_bench_infer2_n1.log 2>&1; grep '^{' gpurun_out/r2_bench_infer2_n1.log | tail -1 | cut -c1-200
for v in strip16 strip64; do echo "== paste $v"; UNIT_B200_LIB=$PWD/unit_b200/build/variants/lib_$v.so timeout 200 python tools/paste_probe.py 2>&1 | tail -1; done
echo "== paste default"; timeout 200 python tools/paste_probe.py 2>&1 | tail -1
