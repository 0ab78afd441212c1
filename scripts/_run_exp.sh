cp tilingnn_b200/_C/libtgnn.so /tmp/keep.so
for m in 6 5; do cp gpurun_exp_m$m.so tilingnn_b200/_C/libtgnn.so; echo "=== MLP warps $m"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_m$m.json 2> gpurun_out/bench_m$m.err; tail -2 gpurun_out/bench_m$m.err
python - gpurun_out/bench_m$m.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "kernel_ms", d["kernel_ms"])
PY
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "persistent or staged or ginw or GINW" 2>&1 | tail -2
done
cp /tmp/keep.so tilingnn_b200/_C/libtgnn.so
