mkdir -p gpurun_out
timeout 300 python tools/micro_roi.py > gpurun_out/micro11.log 2>&1; tail -1 gpurun_out/micro11.log
UNIT_ROI_BWD_SWEEP3=0 timeout 300 python tools/micro_roi.py > gpurun_out/micro11_nosweep3.log 2>&1; tail -1 gpurun_out/micro11_nosweep3.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench7.log 2>&1; tail -1 gpurun_out/bench7.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['roofline']['per_kernel'])"
