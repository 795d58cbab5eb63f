#!/bin/bash
# contexts-per-GPU sweep of the 4K bench workload (device-resident value and e2e)
mkdir -p gpurun_out
for s in 2 3 4 6 8; do
  BENCH_STREAMS=$s timeout 300 python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu > gpurun_out/bench_s$s.log 2> gpurun_out/bench_s$s.err; echo "streams $s rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_s$s.log").read().strip().splitlines()[-1])
print("streams $s value", d["value"], "e2e", d["e2e"]["value"], "pageable", d.get("e2e_pageable",{}).get("value"), "ms/step", d["ms_per_step"])
PY
done
