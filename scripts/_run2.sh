set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "gin_staged or synthetic_configs or determinism or edge_order" 2>&1 | tail -25
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2b_ginw.json 2> $OUT/bench_r2b_ginw.err
tail -2 $OUT/bench_r2b_ginw.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2b_ginw.json")); print("GINW", d["ms_per_step"], d["kernel_ms"])
PY
python tests/diag_eval_bn_stage.py c1_heart.npz c1_complete.npz 2>&1 | tail -50
