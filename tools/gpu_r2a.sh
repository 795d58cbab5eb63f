#!/bin/bash
# gpurun call: full GPU parity suite + colour-seed stage timings (both distance-transform modes)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 --durations=15 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest.log
timeout 300 python tools/colorseed_times.py 1920 1080 > gpurun_out/cs_times_1080.json 2> gpurun_out/cs_times.err; echo "cs1080 rc=$?"; cat gpurun_out/cs_times_1080.json
timeout 300 python tools/colorseed_times.py 3840 2160 > gpurun_out/cs_times_4k.json 2>> gpurun_out/cs_times.err; echo "cs4k rc=$?"; cat gpurun_out/cs_times_4k.json
tail -5 gpurun_out/cs_times.err
