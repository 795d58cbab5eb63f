mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest.log
timeout 300 python tools/k1_matrix.py 2>&1 | grep -E "tile_w=64" | tee gpurun_out/k1_carry.log
