mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit,memory.total --format=csv > gpurun_out/smi.txt 2>&1
nproc >> gpurun_out/smi.txt
python -c "import cv2; print('cv2', cv2.__version__)" >> gpurun_out/smi.txt 2>&1
timeout 120 ./tools/int_peak > gpurun_out/int_peak.json 2> gpurun_out/int_peak.err; echo "int_peak rc=$?"
timeout 600 python tools/debug_stages.py > gpurun_out/debug.log 2>&1; echo "debug rc=$?"
tail -30 gpurun_out/debug.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest.log
timeout 600 python bench.py --steps 2 --warmup 1 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
