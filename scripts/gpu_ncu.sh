#!/bin/bash
# ncu on the B200 box: launch list (my kernels only) + full capture of selected kernels
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-x}; REGEX=${2:-'k_conv_adj|k_gin|k_dense'}; NODES=${3:-1000000}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|k_' -c 300 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --nodes $NODES > $OUT/ncu_launch_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"$REGEX" -s 12 -c 6 \
    -o $OUT/prof_${TAG} -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --nodes $NODES > $OUT/ncu_full_${TAG}.log 2>&1
tail -3 $OUT/ncu_full_${TAG}.log
ls -la $OUT | tail -8
