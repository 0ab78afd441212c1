set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "conv_x or (persistent and x) or (adjacency_kernels and x)" 2>&1 | tee $OUT/x_tests.log | grep -E "passed|failed|rror|max err|max diff"
for k in x h; do
TGNN_CONV=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_x_$k.json 2> $OUT/bench_x_$k.err
python - $OUT/bench_x_$k.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print("failed", sys.argv[1], e)
PY
tail -2 $OUT/bench_x_$k.err
done
