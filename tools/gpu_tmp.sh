mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/shard_large_image.py --size 4096 --verify > gpurun_out/shard2_4096.log 2>&1; echo "shard rc=$?"
tail -2 gpurun_out/shard2_4096.log
