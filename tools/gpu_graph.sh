mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_heads_gpu.py -m gpu -q -x > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_graph.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench6.log 2>&1; tail -3 gpurun_out/bench6.log | cut -c1-1500
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-graph > gpurun_out/bench6_nograph.log 2>&1; tail -1 gpurun_out/bench6_nograph.log | cut -c1-200
