#!/bin/bash
# Round-2 multi-GPU pass 3 (gpurun --gpus 8): final train arm at 8 / 4 / 2, inference arm (bf16 upload) at 8 / 2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-aux --no-cpu-baseline > gpurun_out/r2_bench3_n$n.log 2>&1
  echo "train n=$n rc=$?"; grep '^{' gpurun_out/r2_bench3_n$n.log | tail -1 | python -c "import sys,json;d=json.loads(sys.stdin.read());print(d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],4))"
done
for n in 8 2; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29600 + n)) bench.py --mode infer --gpus $n --steps 50 --warmup 5 > gpurun_out/r2_bench_infer3_n$n.log 2>&1
  echo "infer n=$n rc=$?"; grep '^{' gpurun_out/r2_bench_infer3_n$n.log | tail -1 | python -c "import sys,json;d=json.loads(sys.stdin.read());print(d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],4),'dets',d['detections_last_step'],d.get('detections_gathered'))"
done
