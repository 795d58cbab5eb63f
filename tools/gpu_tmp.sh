mkdir -p gpurun_out
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:meanshift_tile -s 4 -c 2 -o gpurun_out/prof_k1_final python tools/profile_step.py 2 > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_final.csv python tools/profile_step.py 2 > gpurun_out/ncu1.log 2>&1
echo "list rc=$?"
