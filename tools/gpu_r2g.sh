#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi_v2.py tests/test_gpu_sweep.py tests/test_gpu_sharded.py tests/test_gpu_edge_cases.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_merge.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_merge.log
timeout 300 python tools/stage_rooflines.py > gpurun_out/stage_new.log 2>&1; echo "stages rc=$?"; cut -c1-330 gpurun_out/stage_new.log
