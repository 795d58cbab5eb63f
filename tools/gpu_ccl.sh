mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ccl_tiles.py tests/test_gpu_parity.py tests/test_seeds.py tests/test_color_seeds.py -m gpu -q -x --timeout 600 > gpurun_out/pytest_ccl.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_ccl.log
for sz in ${SIZES:-1920x1080 8192x8192}; do
  set -- ${sz%x*} ${sz#*x}
  timeout 300 python tools/profile_stages.py $1 $2 2 > gpurun_out/stages_$1.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/stages_$1.log; }
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_stages_$1.csv python tools/profile_stages.py $1 $2 2 > gpurun_out/ncu_stages_$1.log 2>&1; echo "ncu rc=$?"
  python tools/summarise_launches.py gpurun_out/launches_stages_$1.csv > gpurun_out/launches_stages_$1.md 2>&1; grep "ccl_\|merge_\|scan_\|render\|total" gpurun_out/launches_stages_$1.md
done
if [ -n "$FULL" ]; then
timeout 300 python tools/profile_step.py 1 8192 8192 > gpurun_out/step8192.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$FULL" -c 6 -o gpurun_out/prof_stages_8192 -f python tools/profile_step.py 1 8192 8192 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
fi
