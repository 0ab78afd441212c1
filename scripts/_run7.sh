set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 2>&1 | tail -8
python scripts/role_cycles.py 1000000 32 2>&1 | tail -32
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2g_default.json 2> $OUT/bench_r2g_default.err
tail -2 $OUT/bench_r2g_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2g_default.json")); print("default", d["ms_per_step"], d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print("failed", e)
PY
python bench.py --config5 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config5', d['value'], d['times'], d['network_calls'])"
