mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/shard_large_image.py --size 8192 --verify > gpurun_out/shard8_8192.log 2>&1; echo "shard8k rc=$?"; tail -1 gpurun_out/shard8_8192.log
timeout 600 $TR --master-port 29522 tools/shard_large_image.py --size 16384 > gpurun_out/shard8_16384.log 2>&1; echo "shard16k rc=$?"; tail -1 gpurun_out/shard8_16384.log
timeout 600 $TR --master-port 29523 tools/run_sweep.py --mode batch --frames 256 > gpurun_out/batch8_256.log 2>&1; echo "batch rc=$?"; tail -1 gpurun_out/batch8_256.log
timeout 600 $TR --master-port 29525 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench8.log 2> gpurun_out/bench8.err; echo "bench8 rc=$?"; tail -1 gpurun_out/bench8.log | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29526 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/bench4.log 2> gpurun_out/bench4.err; echo "bench4 rc=$?"; tail -1 gpurun_out/bench4.log | cut -c1-200
