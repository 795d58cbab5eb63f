#!/bin/bash
# gpurun call: colour-seed / distance-transform parity + timings, watershed timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_color_seeds.py tests/test_gpu_fullsize_cv2.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_cs.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_cs.log
timeout 300 python tools/colorseed_times.py 1920 1080 > gpurun_out/cs_times_1080.json 2> gpurun_out/cs_times.err; echo "cs1080 rc=$?"; cat gpurun_out/cs_times_1080.json
timeout 300 python tools/colorseed_times.py 3840 2160 > gpurun_out/cs_times_4k.json 2>> gpurun_out/cs_times.err; echo "cs4k rc=$?"; cat gpurun_out/cs_times_4k.json
tail -5 gpurun_out/cs_times.err
timeout 600 python tools/watershed_times.py 1920 1080 296 > gpurun_out/ws_times_1080.log 2>&1; echo "ws rc=$?"; tail -8 gpurun_out/ws_times_1080.log
