for v in "X=1" "TGNN_DENSE=tf32" "TGNN_GIN=tf32" "TGNN_CONV=chunk" "TGNN_DENSE=ffma"; do
echo "=== $v"
env $v timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 600 -k "config1_eval" 2>&1 | grep -E "passed|failed|rror|c1_heart|c1_complete" | tail -6
done
