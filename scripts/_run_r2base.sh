set -u
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi -L
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tee $OUT/gpu_tests_r2base.log | tail -8
bash scripts/gpu_profile.sh r2base 2>&1 | tail -20
python scripts/small_latency.py > $OUT/small_r2base.log 2>&1; tail -12 $OUT/small_r2base.log
