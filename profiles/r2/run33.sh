mkdir -p gpurun_out
timeout 80 python -m pytest tests/test_ops_gpu.py -q -m gpu -x > gpurun_out/pytest_ops_final3.log 2>&1; echo "ops pytest rc=$?"; tail -2 gpurun_out/pytest_ops_final3.log
