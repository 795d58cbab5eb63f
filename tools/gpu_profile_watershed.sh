#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/profile_watershed.py 960 540 > gpurun_out/ws_plain.log 2>&1; echo "plain rc=$?"; tail -1 gpurun_out/ws_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ws_flood -c 1 -o gpurun_out/prof_ws -f python tools/profile_watershed.py 960 540 > gpurun_out/ncu_ws.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_ws.log
