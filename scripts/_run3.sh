set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "adjacency_kernels or range_guard or edge_block or gin_staged or synthetic_configs" 2>&1 | tail -30
for v in default h; do
  if [ $v = h ]; then export TGNN_CONV=h; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2c_$v.json 2> $OUT/bench_r2c_$v.err
  tail -2 $OUT/bench_r2c_$v.err
  python - $v <<'PY'
import json,sys
d=json.load(open(f"gpurun_out/bench_r2c_{sys.argv[1]}.json")); print(sys.argv[1], d["ms_per_step"], d["kernel_ms"], d["roofline"]["kernel"])
PY
done
