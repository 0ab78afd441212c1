set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s1}
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tee $OUT/gpu_tests_${TAG}.log | tail -6
echo "--- default"; python scripts/small_forward.py 2>&1 | tail -1
python scripts/small_forward.py --eager 2>&1 | tail -1
python scripts/small_forward.py --lattice 10000 8 2>&1 | tail -1
python scripts/small_forward.py --lattice 10000 8 --eager 2>&1 | tail -1
python scripts/small_forward.py --lattice 10000 32 2>&1 | tail -1
echo "--- TGNN_BRANCHES=0"; TGNN_BRANCHES=0 python scripts/small_forward.py 2>&1 | tail -1
TGNN_BRANCHES=0 python scripts/small_forward.py --lattice 10000 8 2>&1 | tail -1
echo "--- TGNN_BNFIN=launch"; TGNN_BNFIN=launch python scripts/small_forward.py 2>&1 | tail -1
TGNN_BNFIN=launch python scripts/small_forward.py --lattice 10000 8 2>&1 | tail -1
