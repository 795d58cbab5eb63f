mkdir -p gpurun_out
python tools/profile_step.py 2 > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:meanshift_tile -s 4 -c 2 -o gpurun_out/prof_k1_persist python tools/profile_step.py 2 > gpurun_out/ncu2.log 2>&1
echo "full rc=$?"
