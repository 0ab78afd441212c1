cp tilingnn_b200/_C/libtgnn.so /tmp/keep.so
cp gpurun_exp_sig.so tilingnn_b200/_C/libtgnn.so
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 600 -k "checkpoint or golden or shipped or tier or config1 or c1" 2>&1 | grep -E "passed|failed|rror|eval|tier|train" | tail -30
cp /tmp/keep.so tilingnn_b200/_C/libtgnn.so
