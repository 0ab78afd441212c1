set -u
N=$1
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L | head -8
for sc in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 10 --warmup 3 --scaling $sc --e2e-steps 3 > $OUT/bench_r2_${N}gpu_$sc.json 2> $OUT/bench_r2_${N}gpu_$sc.err
  tail -2 $OUT/bench_r2_${N}gpu_$sc.err
  python - $sc $N <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench_r2_{sys.argv[2]}gpu_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], sys.argv[2], "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], d["kernel_ms"], "parity", d["parity_max_err"], "e2e", d["e2e"]["value"] if d["e2e"] else None)
except Exception as e: print("failed", e)
PY
done
