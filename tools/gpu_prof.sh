mkdir -p gpurun_out
timeout 300 python tools/prof_step.py > gpurun_out/prof_step.log 2>&1
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench5.log 2>&1; tail -1 gpurun_out/bench5.log
