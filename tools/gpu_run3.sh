mkdir -p gpurun_out
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/ncu1.log 2>&1
echo "launchlist rc=$?"
ncu --set full --clock-control none --import-source on -k regex:meanshift_tile -s 4 -c 2 -o gpurun_out/prof_k1 python tools/profile_step.py 2 > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/plain.log gpurun_out/ncu1.log gpurun_out/ncu2.log
ls -la gpurun_out
