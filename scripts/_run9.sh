set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "adjacency_kernels or persistent_pipelines or gin_staged" 2>&1 | tail -8
for t in 1; do
  export TGNN_CONV_T_TEAMS=$t
  python scripts/role_cycles.py 1000000 32 2>&1 | grep -v "^  warp  [1-7]:" | tail -28
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2i_t$t.json 2> $OUT/bench_r2i_t$t.err
  tail -2 $OUT/bench_r2i_t$t.err
  python - $t <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/bench_r2i_t{sys.argv[1]}.json")); print("teams", sys.argv[1], d["ms_per_step"], d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print("failed", e)
PY
done
