#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4 2; do
  timeout 200 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 20 --warmup 5 --no-aux --no-cpu-baseline > gpurun_out/r2_bench4_n$n.log 2>&1
  echo "train n=$n rc=$?"; grep '^{' gpurun_out/r2_bench4_n$n.log | tail -1 | python -c "import sys,json;d=json.loads(sys.stdin.read());print(d['n_gpus'],'value',round(d['value']),'ms',round(d['ms_per_step'],4),'e2e',round(d['e2e']['value']),round(d['e2e']['ms_per_step'],4))"
done
