mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/bench_n8.log 2>&1; tail -1 gpurun_out/bench_n8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches','cuda_graphs')})" || tail -30 gpurun_out/bench_n8.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4.log 2>&1; tail -1 gpurun_out/bench_n4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e')})" || tail -30 gpurun_out/bench_n4.log
