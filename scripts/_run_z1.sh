set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "(adjacency_kernels and z) or conv_z or (persistent and z) or (range_guard and z)" 2>&1 | tee $OUT/z5_tests.log | tail -25
summ() {
python - "$1" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"], d["roofline"]["kernel"], "parity", d.get("parity_max_err"))
except Exception as e: print("failed", sys.argv[1], e)
PY
}
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_z5.json 2> $OUT/bench_z5.err
summ $OUT/bench_z5.json; tail -3 $OUT/bench_z5.err
TGNN_ROLE_DBG=1 timeout 300 python scripts/role_cycles.py > $OUT/z5_roles.log 2>&1; tail -30 $OUT/z5_roles.log
