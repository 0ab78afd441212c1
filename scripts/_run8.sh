set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s --timeout 300 -k "adjacency_kernels or range_guard or persistent_pipelines" 2>&1 | tail -12
python scripts/role_cycles.py 1000000 32 2>&1 | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_r2h_default.json 2> $OUT/bench_r2h_default.err
tail -2 $OUT/bench_r2h_default.err
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/bench_r2h_default.json")); print("default", d["ms_per_step"], d["kernel_ms"], d["roofline"]["kernel"])
except Exception as e: print("failed", e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_gin_w|k_conv_t$' -s 4 -c 2 -o $OUT/prof_r2h -f python scripts/role_cycles.py 1000000 32 > $OUT/ncu_r2h.log 2>&1
tail -3 $OUT/ncu_r2h.log; ls -la $OUT/prof_r2h* 
