set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s4}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tee $OUT/gpu_tests_${TAG}.log | tail -4
echo "--- default"; python scripts/small_forward.py 2>&1 | tail -1
python scripts/small_forward.py --eager 2>&1 | tail -1
python scripts/small_forward.py --lattice 10000 8 2>&1 | tail -1
python scripts/small_forward.py --lattice 10000 32 2>&1 | tail -1
python scripts/small_forward.py --lattice 3000 8 2>&1 | tail -1
echo "--- TGNN_CONV_CLUSTER=0"; TGNN_CONV_CLUSTER=0 python scripts/small_forward.py 2>&1 | tail -1
TGNN_CONV_CLUSTER=0 python scripts/small_forward.py --lattice 3000 8 2>&1 | tail -1
bash scripts/_run_dbg.sh 2>&1 | grep "warp" | head -8
