mkdir -p gpurun_out
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches2.csv python tools/profile_step.py 2 > gpurun_out/ncu1.log 2>&1
echo "launchlist rc=$?"
