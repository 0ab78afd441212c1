#!/bin/bash
# quick GPU check: parity tests + headline bench (no ncu)
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-q}
timeout 900 python -m pytest tests -m gpu -x -q -s --timeout 300 2>&1 | tail -25
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
python - <<PY
import json
d=json.load(open("$OUT/bench_${TAG}.json"))
print("value", d["value"], "ms", d["ms_per_step"], "kernel_ms", d["kernel_ms"], "e2e", d["e2e"]["value"] if d["e2e"] else None, "roof", d["roofline"]["frac"], d["roofline"]["forward"]["frac"])
PY
tail -3 $OUT/bench_${TAG}.err
