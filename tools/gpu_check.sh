#!/bin/bash
# One gpurun call: GPU parity tests, smoke, short bench (ours + reference arm), ncu launch list.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python -X faulthandler bench.py --steps ${STEPS:-5} --warmup ${WARMUP:-3} > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
if [ -z "$NO_REF" ]; then
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
fi
if [ -z "$NO_NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_4k.csv python tools/profile_step.py 2 > gpurun_out/ncu_4k.log 2>&1; echo "ncu rc=$?"
python tools/summarise_launches.py gpurun_out/launches_4k.csv > gpurun_out/launches_4k.md 2>&1; tail -30 gpurun_out/launches_4k.md
fi
