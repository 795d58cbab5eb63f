mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q --timeout 600 > gpurun_out/pytest_sharded.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_sharded.log
