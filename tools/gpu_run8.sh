mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log
timeout 600 python tools/k1_matrix.py > gpurun_out/k1_matrix.log 2>&1; echo "matrix rc=$?"; cat gpurun_out/k1_matrix.log
timeout 900 python -X faulthandler bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
