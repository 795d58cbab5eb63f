#!/bin/bash
# gpurun call: sharded tests with the derived halo bound, ncu --set full of the float distance-transform kernel, big watershed batch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_shard.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/pytest_shard.log
timeout 200 python tools/profile_colorseeds.py 1920 1080 > gpurun_out/cs_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dt_wave2 -c 1 -o gpurun_out/prof_dt2 -f python tools/profile_colorseeds.py 1920 1080 > gpurun_out/ncu_dt2.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_dt2.log
timeout 900 python tools/watershed_times.py 1920 1080 1184 > gpurun_out/ws_times_1080_big.log 2>&1; echo "ws rc=$?"; tail -4 gpurun_out/ws_times_1080_big.log | cut -c1-400
