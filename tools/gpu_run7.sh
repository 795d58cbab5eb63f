mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_sweep.py tests/test_cli.py -m gpu -q --timeout 900 > gpurun_out/pytest_sweep.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_sweep.log
timeout 300 python tools/run_sweep.py --mode sweep > gpurun_out/sweep1.log 2>&1; echo "sweep rc=$?"; tail -2 gpurun_out/sweep1.log
