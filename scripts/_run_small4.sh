set -u
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s3}
echo "--- default"; python scripts/small_forward.py 2>&1 | tail -1
python scripts/small_forward.py --lattice 10000 8 2>&1 | tail -1
python scripts/small_forward.py --lattice 10000 32 2>&1 | tail -1
echo "--- TGNN_CONV_PREF=0"; TGNN_CONV_PREF=0 python scripts/small_forward.py 2>&1 | tail -1
summ() {
python - "$@" <<'PY'
import csv, sys, collections
for f in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); ui = hdr.index("Metric Unit")
    rows = rows[1:]
    names = [r[ki] for r in rows]
    starts = [i for i, n in enumerate(names) if "k_init<0>" in n]
    s = starts[-1]
    agg = collections.OrderedDict(); tot = 0.0
    for r in rows[s:]:
        v = float(r[vi].replace(",", "")); v = v / 1000.0 if r[ui] in ("ns", "nsecond") else v
        k = r[ki].split("(")[0][-40:]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
    print(f, "last forward: %d launches, sum of kernel durations %.1f us" % (len(rows) - s, tot))
    for k, (c, v) in agg.items(): print("   %-42s x%-3d %8.1f us  (%.2f us each)" % (k, c, v, v / c))
PY
}
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_' -c 900 --csv --log-file $OUT/launches_small_warm_${TAG}.csv \
    python scripts/small_forward.py --eager --reps 2 > $OUT/ncu_small_${TAG}.log 2>&1
summ $OUT/launches_small_warm_${TAG}.csv
TGNN_CONV_PREF=0 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_conv' -c 200 --csv --log-file $OUT/launches_small_warm_nopref_${TAG}.csv \
    python scripts/small_forward.py --eager --reps 2 > $OUT/ncu_small_${TAG}.log 2>&1
grep k_conv_h $OUT/launches_small_warm_nopref_${TAG}.csv | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:'k_' -c 500 --csv --log-file $OUT/launches_10k_warm_${TAG}.csv \
    python scripts/small_forward.py --eager --reps 2 --lattice 10000 8 > $OUT/ncu_10k_${TAG}.log 2>&1
summ $OUT/launches_10k_warm_${TAG}.csv
