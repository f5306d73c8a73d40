mkdir -p gpurun_out
timeout 100 ncu --metrics gpu__time_duration.sum -c 1 env > gpurun_out/ncu_env.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r1c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_list3.log 2>&1; tail -2 gpurun_out/ncu_list3.log | cut -c1-300
