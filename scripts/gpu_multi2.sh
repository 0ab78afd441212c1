set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 1500 python -m pytest tests/test_gpu_shard.py -m gpu -x -q -s --timeout 600 2>&1 | tee $OUT/shard_tests_2gpu_r2.log | tail -15
for sc in weak strong; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 --scaling $sc --e2e-steps 2 > $OUT/bench_r2_2gpu_$sc.json 2> $OUT/bench_r2_2gpu_$sc.err
  tail -3 $OUT/bench_r2_2gpu_$sc.err
  python - $sc <<'PY'
import json,sys
try:
    d=json.loads(open(f"gpurun_out/bench_r2_2gpu_{sys.argv[1]}.json").read().strip().splitlines()[-1]); print(sys.argv[1], d["value"], d["ms_per_step"], d["kernel_ms"], "parity", d["parity_max_err"], "e2e", d["e2e"]["value"] if d["e2e"] else None)
except Exception as e: print("failed", e)
PY
done
