mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest11.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest11.log
