#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_color_seeds.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_cs.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_cs.log
timeout 300 python tools/colorseed_times.py 1920 1080 > gpurun_out/cs_times_1080.json 2> gpurun_out/cs_times.err; echo "cs1080 rc=$?"; cut -c1-700 gpurun_out/cs_times_1080.json
timeout 300 python tools/colorseed_times.py 3840 2160 > gpurun_out/cs_times_4k.json 2>> gpurun_out/cs_times.err; echo "cs4k rc=$?"; cut -c1-700 gpurun_out/cs_times_4k.json
tail -5 gpurun_out/cs_times.err
if [ -n "$NCU" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dt_wave2 -c 1 -o gpurun_out/prof_dt2b -f python tools/profile_colorseeds.py 1920 1080 > gpurun_out/ncu_dt2b.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dt2b.log
fi
