set -u
OUT=gpurun_out; mkdir -p $OUT
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_score.json 2> $OUT/bench_score.err; tail -2 $OUT/bench_score.err
python - $OUT/bench_score.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"])
PY
cp tilingnn_b200/_C/libtgnn.so /tmp/keep.so; cp gpurun_exp_dense.so tilingnn_b200/_C/libtgnn.so
TGNN_DENSE_DBG=1 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_densedbg.json 2> $OUT/bench_densedbg.err
grep "dbg warp" $OUT/bench_densedbg.err | head -70
cp /tmp/keep.so tilingnn_b200/_C/libtgnn.so
