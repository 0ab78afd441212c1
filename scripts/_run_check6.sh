set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-c6}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 600 -k "small_graph" 2>&1 | grep -E "passed|failed|rror|max err|max diff" | tail -30
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err; tail -2 $OUT/bench_${TAG}.err
python - $OUT/bench_${TAG}.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
PY
python bench.py --config5 --steps 20 --warmup 3 > $OUT/bench_${TAG}_config5.json 2>> $OUT/bench_${TAG}.err; cat $OUT/bench_${TAG}_config5.json | cut -c1-700
python scripts/small_latency.py 2>&1 | tail -7
