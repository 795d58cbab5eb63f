#!/bin/bash
# 8-GPU call: strip sharding through the library path (verify vs unsharded, vs the CPU oracle, label-only timing), the 8-GPU bench line
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/shard_large_image.py --size 8192 --verify > gpurun_out/shard8_8192_verify.log 2>&1; echo "verify rc=$?"; tail -1 gpurun_out/shard8_8192_verify.log | cut -c1-1200
timeout 900 $TR --master-port 29522 tests/shard_verify_oracle.py --size 16384 > gpurun_out/shard8_16384_oracle.log 2>&1; echo "oracle rc=$?"; tail -1 gpurun_out/shard8_16384_oracle.log | cut -c1-1400
timeout 600 $TR --master-port 29523 tools/shard_large_image.py --size 16384 --min-size 0 --color-dist 0 > gpurun_out/shard8_16384_nomerge.log 2>&1; echo "nomerge rc=$?"; tail -1 gpurun_out/shard8_16384_nomerge.log | cut -c1-900
timeout 600 $TR --master-port 29524 tools/shard_large_image.py --size 16384 > gpurun_out/shard8_16384_merge.log 2>&1; echo "merge rc=$?"; tail -1 gpurun_out/shard8_16384_merge.log | cut -c1-900
timeout 600 $TR --master-port 29525 bench.py --gpus 8 --steps 5 --warmup 3 --no-secondary > gpurun_out/bench8.log 2> gpurun_out/bench8.err; echo "bench8 rc=$?"; tail -1 gpurun_out/bench8.log | cut -c1-2500
