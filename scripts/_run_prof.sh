set -u
bash scripts/gpu_profile.sh r2 2>&1 | tail -12
python scripts/small_latency.py > gpurun_out/small_r2.log 2>&1; tail -8 gpurun_out/small_r2.log
