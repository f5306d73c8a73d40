mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest12.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest12.log
