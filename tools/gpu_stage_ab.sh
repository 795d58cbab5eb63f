#!/bin/bash
# One gpurun call for the label / merge stage kernels: GPU parity suite with the defaults, stage rooflines and the step-level
# overlap probe (AB="0 1": also with ccl_quad = 0, merge_strips = 0, the kernels before), launch lists at 4K and 8192^2,
# FULL=regex: ncu --set full of the matching kernels at 8192^2.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/pytest_stage_ab.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_stage_ab.log
for v in ${AB:-1}; do
  MSG_CCL_QUAD=$v MSG_MERGE_STRIPS=$v timeout 600 python tools/stage_rooflines.py > gpurun_out/stage_ab_$v.log 2>&1; echo "stages $v rc=$?"; grep '"size"' gpurun_out/stage_ab_$v.log | cut -c1-330
  MSG_CCL_QUAD=$v MSG_MERGE_STRIPS=$v timeout 300 python tools/overlap_probe.py > gpurun_out/stage_ab_probe_$v.log 2>&1; echo "probe $v rc=$?"; head -3 gpurun_out/stage_ab_probe_$v.log
done
for sz in 3840x2160 8192x8192; do
  set -- ${sz%x*} ${sz#*x}
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_stages_$1.csv python tools/profile_stages.py $1 $2 2 > gpurun_out/ncu_stages_$1.log 2>&1; echo "ncu rc=$?"
  python tools/summarise_launches.py gpurun_out/launches_stages_$1.csv > gpurun_out/launches_stages_$1.md 2>&1; grep "ccl_\|merge_\|scan_\|render\|total" gpurun_out/launches_stages_$1.md
done
if [ -n "$FULL" ]; then
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:"$FULL" -c 3 -o gpurun_out/prof_stages_8192 -f python tools/profile_stages.py 8192 8192 1 > gpurun_out/ncu_full_stages.log 2>&1; echo "ncu full rc=$?"; tail -2 gpurun_out/ncu_full_stages.log
fi
