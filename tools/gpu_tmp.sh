mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | cut -c1-120
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_inst.csv python tools/profile_step.py 2 > gpurun_out/ncu1.log 2>&1
echo "list rc=$?"
