cp tilingnn_b200/_C/libtgnn.so /tmp/keep.so
for e in 4 5 6; do cp gpurun_exp_$e.so tilingnn_b200/_C/libtgnn.so; echo "=== EXP $e"; bash scripts/_run_dbg.sh 2>&1 | grep "warp  [013]"; done
cp /tmp/keep.so tilingnn_b200/_C/libtgnn.so
