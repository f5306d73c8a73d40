mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest10.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest10.log
timeout 300 python tools/micro_roi.py > gpurun_out/micro14.log 2>&1; tail -1 gpurun_out/micro14.log
timeout 300 python tools/micro_roi.py --bf16 > gpurun_out/micro14_bf16.log 2>&1; tail -1 gpurun_out/micro14_bf16.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench8.log 2>&1; tail -1 gpurun_out/bench8.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['roofline']['per_kernel'])"
