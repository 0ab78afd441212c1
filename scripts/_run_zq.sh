set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=$1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "(adjacency_kernels and z) or conv_z or (persistent and z)" 2>&1 | tee $OUT/${TAG}_tests.log | grep -E "passed|failed|rror|z:|z32:|multi|300k"
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
python - $OUT/bench_${TAG}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print("failed", sys.argv[1], e)
PY
tail -2 $OUT/bench_${TAG}.err
