mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/shard_large_image.py --size 8192 --verify > gpurun_out/shard8_8192.log 2>&1; echo "shard8k rc=$?"; tail -1 gpurun_out/shard8_8192.log
timeout 600 $TR --master-port 29522 tools/shard_large_image.py --size 16384 > gpurun_out/shard8_16384.log 2>&1; echo "shard16k rc=$?"; tail -1 gpurun_out/shard8_16384.log
timeout 600 $TR --master-port 29523 tools/run_sweep.py --mode batch --frames 256 > gpurun_out/batch8_256.log 2>&1; echo "batch rc=$?"; tail -1 gpurun_out/batch8_256.log
timeout 600 $TR --master-port 29524 tools/run_sweep.py --mode sweep > gpurun_out/sweep8.log 2>&1; echo "sweep rc=$?"; tail -1 gpurun_out/sweep8.log
timeout 600 $TR --master-port 29525 bench.py --gpus 8 --steps 3 --warmup 2 > gpurun_out/bench8.log 2> gpurun_out/bench8.err; echo "bench8 rc=$?"; tail -1 gpurun_out/bench8.log | cut -c1-400
