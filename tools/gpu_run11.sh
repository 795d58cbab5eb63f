mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_edge_cases.py -m gpu -q --timeout 600 > gpurun_out/pytest_edge.log 2>&1; echo "pytest rc=$?"
tail -40 gpurun_out/pytest_edge.log
