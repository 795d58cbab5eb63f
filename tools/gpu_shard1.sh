mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_fullsize_cv2.py -m gpu -q -x --timeout 800 > gpurun_out/pytest_shard.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_shard.log
