#!/bin/bash
# 2-GPU call: strip sharding through the library path, sharded == unsharded (4096^2) and == CPU oracle
mkdir -p gpurun_out
N=2
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29531 tools/shard_large_image.py --size 4096 --verify > gpurun_out/shard2_verify.log 2>&1; echo "verify rc=$?"; tail -1 gpurun_out/shard2_verify.log | cut -c1-1000
timeout 600 $TR --master-port 29532 tests/shard_verify_oracle.py --size 4096 > gpurun_out/shard2_oracle.log 2>&1; echo "oracle rc=$?"; tail -1 gpurun_out/shard2_oracle.log | cut -c1-1000
