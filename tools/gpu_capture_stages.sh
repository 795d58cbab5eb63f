mkdir -p gpurun_out
for sz in 1920x1080 3840x2160 8192x8192; do
  set -- ${sz%x*} ${sz#*x}
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 300 --csv \
    --log-file gpurun_out/launches_stages_$1.csv python tools/profile_stages.py $1 $2 2 > gpurun_out/ncu_stages_$1.log 2>&1; echo "ncu rc=$?"
  python tools/summarise_launches.py gpurun_out/launches_stages_$1.csv > gpurun_out/launches_stages_$1.md 2>&1; grep "total" gpurun_out/launches_stages_$1.md
done
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"ccl_tile4|merge_stats|ccl_border" -c 3 -o gpurun_out/prof_stages_8192 -f python tools/profile_stages.py 8192 8192 1 > gpurun_out/ncu_full_stages.log 2>&1; echo "ncu full rc=$?"
