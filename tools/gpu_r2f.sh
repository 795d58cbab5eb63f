#!/bin/bash
# merge grid A/B: label / merge / render stage timings at 4K and 8192^2 for several cooperative-grid sizes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_abi_v2.py tests/test_gpu_sweep.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_merge.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_merge.log
for g in 0 16 32 64 148 296; do
  SIZES=3840x2160,8192x8192 MERGE_GRID=$g timeout 300 python tools/stage_rooflines.py > gpurun_out/stage_grid_$g.log 2>&1; echo "grid $g rc=$?"
  python - <<PY
import json
for l in open("gpurun_out/stage_grid_$g.log"):
    if l.startswith("{"):
        r=json.loads(l); print("grid $g", r["size"], "rounds", r["merge_rounds"], "merge ms", r["merge"]["ms"], "label ms", r["label"]["ms"])
PY
done
