set -u
OUT=gpurun_out; mkdir -p $OUT
python scripts/role_cycles.py 1000000 32 2>&1 | tail -40
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2e_default.json 2> $OUT/bench_r2e_default.err
tail -2 $OUT/bench_r2e_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2e_default.json")); print("default", d["ms_per_step"], d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print("failed", e)
PY
