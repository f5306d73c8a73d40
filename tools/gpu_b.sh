mkdir -p gpurun_out
timeout 400 python bench.py --steps 30 --warmup 5 > gpurun_out/bench9.log 2>&1; tail -1 gpurun_out/bench9.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['roofline']['per_kernel']); print(d['roofline']['traffic'], d['aux']); print(d['cpu_baseline'])" || tail -20 gpurun_out/bench9.log
