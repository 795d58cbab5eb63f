#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_jni_glue_runtime.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_jni.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_jni.log
timeout 120 python tools/sanitize_small.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/sanitize_plain.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -12 gpurun_out/sanitize_memcheck.log | cut -c1-300
