set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "persistent_pipelines or gin_staged or graph_replay" 2>&1 | tail -20
for v in default h hw0; do
  unset TGNN_CONV TGNN_GINW
  if [ $v = h ]; then export TGNN_CONV=h; fi
  if [ $v = hw0 ]; then export TGNN_CONV=h TGNN_GINW=0; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2d_$v.json 2> $OUT/bench_r2d_$v.err
  tail -2 $OUT/bench_r2d_$v.err
  python - $v <<'PY'
import json,sys
try:
    d=json.load(open(f"gpurun_out/bench_r2d_{sys.argv[1]}.json")); print(sys.argv[1], d["ms_per_step"], d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print(sys.argv[1], "failed", e)
PY
done
unset TGNN_CONV TGNN_GINW
python scripts/small_latency.py 2>&1 | tail -8
python bench.py --steps 20 --warmup 5 --nodes 10000 --deg 8 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('10k', d['ms_per_step'], d['kernel_ms'])"
