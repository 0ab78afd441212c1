set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_greedy.py -m gpu -x -q -s --timeout 300 -k "adjacency_kernels or range_guard or edge_block or gin_staged or persistent_pipelines or graph_replay or node_mask or config5" 2>&1 | tail -30
python scripts/role_cycles.py 1000000 32 2>&1 | tail -32
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2f_default.json 2> $OUT/bench_r2f_default.err
tail -2 $OUT/bench_r2f_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2f_default.json")); print("default", d["ms_per_step"], d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print("failed", e)
PY
python scripts/small_latency.py 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --nodes 10000 --deg 8 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('10k', d['ms_per_step'], d['kernel_ms'])"
python bench.py --config5 --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('config5', d['value'], d['times'], d['network_calls'])"
