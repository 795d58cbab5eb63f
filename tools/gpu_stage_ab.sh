#!/bin/bash
# One gpurun call for the label / merge stage kernels: the labelling / merge parity tests, then stage rooflines and the
# step-level overlap probe for every "ccl_quad:merge_strips" setting in AB (default: the library defaults), launch lists.
mkdir -p gpurun_out
MSG_CCL_QUAD=${TQ:-1} MSG_MERGE_STRIPS=${TM:-1} timeout 900 python -m pytest tests/test_gpu_ccl_tiles.py tests/test_gpu_parity.py tests/test_gpu_sharded.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_stage_ab.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_stage_ab.log
for v in ${AB:-1:1}; do
  q=${v%:*}; m=${v#*:}
  MSG_CCL_QUAD=$q MSG_MERGE_STRIPS=$m timeout 600 python tools/stage_rooflines.py > gpurun_out/stage_ab_${q}_$m.log 2>&1; echo "stages quad=$q strips=$m rc=$?"; grep '"size"' gpurun_out/stage_ab_${q}_$m.log | cut -c1-330
  MSG_CCL_QUAD=$q MSG_MERGE_STRIPS=$m timeout 300 python tools/overlap_probe.py > gpurun_out/stage_ab_probe_${q}_$m.log 2>&1; echo "probe rc=$?"; head -3 gpurun_out/stage_ab_probe_${q}_$m.log
done
