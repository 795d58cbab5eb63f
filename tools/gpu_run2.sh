mkdir -p gpurun_out
timeout 900 python -X faulthandler bench.py --steps 3 --warmup 2 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -5 gpurun_out/bench.log; tail -30 gpurun_out/bench.err
