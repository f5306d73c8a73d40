mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_heads_gpu.py -m gpu -q > gpurun_out/pytest_heads.log 2>&1; tail -2 gpurun_out/pytest_heads.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.log 2>&1; tail -1 gpurun_out/bench_n2.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches','cuda_graphs')})" || tail -20 gpurun_out/bench_n2.log
