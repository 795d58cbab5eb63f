mkdir -p gpurun_out
timeout 600 python tools/k1_matrix.py > gpurun_out/k1_matrix2.log 2>&1; echo "matrix rc=$?"; grep -E "tile_w=64 acc=0|acc=1 : L0 0.*tile_w=64" gpurun_out/k1_matrix2.log; cat gpurun_out/k1_matrix2.log | grep "tile_w=64"
