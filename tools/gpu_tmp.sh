mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q --timeout 900 -k extreme > gpurun_out/pytest_ext.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_ext.log
