set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=$1
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tee $OUT/gpu_tests_${TAG}.log | tail -6
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
python - $OUT/bench_${TAG}.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"], d["roofline"]["kernel"], "frac", round(d["roofline"]["frac"],4))
    print("e2e", d["e2e"])
    print("cpu", d["cpu_baseline"])
except Exception as e: print("failed", sys.argv[1], e)
PY
tail -3 $OUT/bench_${TAG}.err
