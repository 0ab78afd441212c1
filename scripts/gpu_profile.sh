#!/bin/bash
# Run on the B200 box (under gpurun): headline bench, the other configs, ncu launch list, ncu --set full on the top kernels.
# Outputs land in gpurun_out/ ; the summaries that matter are copied to profiles/ afterwards (scripts/ncu_summary.py).
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-r1}
python bench.py --steps 20 --warmup 5 > $OUT/bench_${TAG}.json 2> $OUT/bench_${TAG}.err
tail -c 1500 $OUT/bench_${TAG}.json; tail -3 $OUT/bench_${TAG}.err
for cfg in "10000 8" "10000 32" "100000 32" "100000 8" "1000000 8"; do
  set -- $cfg
  python bench.py --steps 20 --warmup 3 --nodes $1 --deg $2 --no-cpu-baseline --e2e-steps 2 > $OUT/bench_${TAG}_n$1_d$2.json 2>> $OUT/bench_${TAG}.err
done
python bench.py --steps 10 --warmup 3 --bn eval --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_evalbn.json 2>> $OUT/bench_${TAG}.err
TGNN_CONV=s python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_convs.json 2>> $OUT/bench_${TAG}.err
TGNN_CONV=z python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_convz.json 2>> $OUT/bench_${TAG}.err
TGNN_CONV=t python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_convt.json 2>> $OUT/bench_${TAG}.err
TGNN_GINW=0 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_gin.json 2>> $OUT/bench_${TAG}.err
TGNN_DENSE=tf32 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_densetf32.json 2>> $OUT/bench_${TAG}.err
TGNN_CONV=chunk python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $OUT/bench_${TAG}_convtf32.json 2>> $OUT/bench_${TAG}.err
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_${TAG}_reference.json 2>> $OUT/bench_${TAG}.err
python bench.py --config5 --steps 20 --warmup 3 > $OUT/bench_${TAG}_config5.json 2>> $OUT/bench_${TAG}.err
# every launch of this library's kernels with its device time (cold cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_' -c 300 --csv --log-file $OUT/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_launch_${TAG}.log 2>&1
# full capture of the heaviest kernels (skip the warm-up forwards' launches)
ncu --set full --clock-control none --import-source on -k regex:'k_conv_h|k_gin|k_dense_tc|k_init|k_combine' -s 24 -c 22 \
    -o $OUT/prof_${TAG} -f python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $OUT/ncu_full_${TAG}.log 2>&1
# real-size layouts: forward latency (captured replay and eager), predict() breakdown, warm-cache per-kernel durations
python scripts/small_forward.py > $OUT/small_${TAG}.log 2>&1
python scripts/small_forward.py --eager >> $OUT/small_${TAG}.log 2>&1
python scripts/small_forward.py --lattice 10000 8 >> $OUT/small_${TAG}.log 2>&1
python scripts/small_forward.py --lattice 10000 32 >> $OUT/small_${TAG}.log 2>&1
python scripts/small_latency.py >> $OUT/small_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_' -c 900 --csv --log-file $OUT/launches_small_warm_${TAG}.csv \
    python scripts/small_forward.py --eager --reps 2 > $OUT/ncu_small_${TAG}.log 2>&1
python scripts/role_cycles.py 1000000 32 > $OUT/role_cycles_${TAG}.log 2>&1
cat $OUT/small_${TAG}.log
ls -la $OUT | tail -14
