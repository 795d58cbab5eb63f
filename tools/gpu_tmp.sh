mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest.log
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches3.csv python tools/profile_step.py 2 > gpurun_out/ncu1.log 2>&1
echo "launchlist rc=$?"
