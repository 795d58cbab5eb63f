#!/bin/bash
# One gpurun call: full GPU parity suite, smoke, stage rooflines, colour-seed timings, 1-GPU bench (ours + reference arm), ncu launch list of the bench step.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python tools/stage_rooflines.py > gpurun_out/stage_new.log 2>&1; echo "stages rc=$?"; tail -4 gpurun_out/stage_new.log
timeout 300 python tools/colorseed_times.py 1920 1080 > gpurun_out/cs_times_1080.json 2> gpurun_out/cs_times.err; cut -c1-420 gpurun_out/cs_times_1080.json
timeout 300 python tools/colorseed_times.py 3840 2160 > gpurun_out/cs_times_4k.json 2>> gpurun_out/cs_times.err; cut -c1-420 gpurun_out/cs_times_4k.json
timeout 900 python -X faulthandler bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -1 gpurun_out/bench.log | cut -c1-1500; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench_ref rc=$?"
tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
  --log-file gpurun_out/launches_4k.csv python tools/profile_step.py 2 > gpurun_out/ncu_4k.log 2>&1; echo "ncu rc=$?"
python tools/summarise_launches.py gpurun_out/launches_4k.csv > gpurun_out/launches_4k.md 2>&1; tail -32 gpurun_out/launches_4k.md
