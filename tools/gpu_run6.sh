mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/shard_large_image.py --size 4096 --verify > gpurun_out/shard2_4096.log 2>&1; echo "shard rc=$?"
tail -3 gpurun_out/shard2_4096.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/shard_large_image.py --size 16384 > gpurun_out/shard2_16384.log 2>&1; echo "shard16k rc=$?"
tail -2 gpurun_out/shard2_16384.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 2 > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "bench2 rc=$?"
tail -2 gpurun_out/bench2.log; tail -5 gpurun_out/bench2.err
