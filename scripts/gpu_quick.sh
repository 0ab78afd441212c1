#!/bin/bash
# quick GPU check: parity tests + headline bench (no ncu).  usage: gpu_quick.sh TAG [extra TGNN_CONV values to A/B]
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-q}; shift || true
timeout 900 python -m pytest tests -m gpu -x -q -s --timeout 300 2>&1 | tail -30
summ() {
python - "$1" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print(sys.argv[1], "value %.3g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"],
      "e2e", d["e2e"]["value"] if d.get("e2e") else None, "roof", round(d["roofline"]["frac"], 4), round(d["roofline"]["forward"]["frac"], 4))
PY
}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
summ $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
for k in "$@"; do
  TGNN_CONV=$k python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_$k.json 2> $OUT/bench_${TAG}_$k.err
  summ $OUT/bench_${TAG}_$k.json; tail -3 $OUT/bench_${TAG}_$k.err
done
