mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_filters.py -m gpu -q --timeout 600 > gpurun_out/pytest_filters.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_filters.log
