set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-c10}
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 600 -k "checkpoint or golden or shipped or tier or config1 or c1" 2>&1 | grep -E "passed|failed|rror|eval|tier|train" | tail -30
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -2 $OUT/bench_${TAG}.err
python - $OUT/bench_${TAG}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"])
PY
