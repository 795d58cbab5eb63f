# multi-GPU checks of the strip pipeline (config 5); N = number of GPUs of the gpurun call
mkdir -p gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29521 tools/shard_large_image.py --size ${SIZE_V:-4096} --verify > gpurun_out/shard${N}_verify.log 2>&1; echo "verify rc=$?"; tail -1 gpurun_out/shard${N}_verify.log | cut -c1-1200
timeout 900 $TR --master-port 29522 tests/shard_verify_oracle.py --size ${SIZE_O:-4096} > gpurun_out/shard${N}_oracle.log 2>&1; echo "oracle rc=$?"; tail -1 gpurun_out/shard${N}_oracle.log | cut -c1-1400
if [ -n "$SIZE_BIG" ]; then
timeout 600 $TR --master-port 29523 tools/shard_large_image.py --size $SIZE_BIG > gpurun_out/shard${N}_${SIZE_BIG}.log 2>&1; echo "big rc=$?"; tail -1 gpurun_out/shard${N}_${SIZE_BIG}.log | cut -c1-1200
fi
