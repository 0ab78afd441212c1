#!/bin/bash
# compute-sanitizer passes over the scoring path: memcheck + racecheck + synccheck on the 10k x deg 8
# configuration (every kernel family runs, incl. the tcgen05 dense stages) and on a small real-layout-sized graph
# (split-tile conv, BatchNorm finish inside k_combine); with 2+ GPUs also memcheck of a sharded forward through the
# peer-memory exchange (epoch flags: st.release.sys / ld.acquire.sys, double buffering by epoch parity).
# Run on the B200 box:  bash scripts/gpu_sanitize.sh [tag]      logs -> gpurun_out/sanitize_<tag>_*.log
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r2}
CS=/usr/local/cuda/bin/compute-sanitizer
DRV='python scripts/sanitize_driver.py'
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 3 $DRV --nodes 10000 --deg 8 --depth 6 \
      > $OUT/sanitize_${TAG}_${tool}_10k.log 2>&1
  echo "$tool 10k: exit $?"; tail -3 $OUT/sanitize_${TAG}_${tool}_10k.log
done
# real-layout size: cluster-split k_conv_h (distributed-shared-memory sum of the partial tiles), k_gin_s, BatchNorm statistics
# finished in the consumers' prologues, the collision branch on a side stream
for tool in memcheck racecheck synccheck; do
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 3 $DRV --nodes 600 --deg 16 --depth 20 \
      > $OUT/sanitize_${TAG}_${tool}_600.log 2>&1
  echo "$tool 600: exit $?"; tail -3 $OUT/sanitize_${TAG}_${tool}_600.log
done
timeout 900 $CS --tool racecheck --print-limit 20 --error-exitcode 3 $DRV --nodes 2500 --deg 8 --depth 6 \
    > $OUT/sanitize_${TAG}_racecheck_2500.log 2>&1
echo "racecheck 2500 (clusters of 2): exit $?"; tail -3 $OUT/sanitize_${TAG}_racecheck_2500.log
# the windowed tcgen05 kernel (opt-in) and the staged-window collision kernel on a graph large enough for both.  memcheck
# only: racecheck does not model mbarrier / async-proxy ordering and slows these pipelines past their bounded waits
# (profiles/r2/sanitize/README.txt)
for tool in memcheck; do
  TGNN_CONV=z TGNN_GINW=1 timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 3 $DRV --nodes 40000 --deg 16 --depth 3 \
      > $OUT/sanitize_${TAG}_${tool}_convz_ginw.log 2>&1
  echo "$tool conv_z + gin_w: exit $?"; tail -3 $OUT/sanitize_${TAG}_${tool}_convz_ginw.log
done
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 1200 $CS --tool memcheck --print-limit 20 --error-exitcode 3 --target-processes all \
      python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29577 \
      scripts/sanitize_driver.py --nodes 20000 --deg 8 --depth 6 > $OUT/sanitize_${TAG}_memcheck_2gpu.log 2>&1
  echo "memcheck 2gpu: exit $?"; tail -3 $OUT/sanitize_${TAG}_memcheck_2gpu.log
fi
