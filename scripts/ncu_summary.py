"""Summarise ncu outputs (run here, no GPU): launch list CSV -> per-kernel totals; .ncu-rep -> key raw metrics."""
import collections, csv, subprocess, sys, json

def launches(path):
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[h]; kn, mv, mn = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[h + 1:]:
        if len(r) <= mv or r[mn] != "gpu__time_duration.sum": continue
        name = r[kn].split("(")[0].replace("void ", "").replace("tgnn::<unnamed>::", "")
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':45s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'ms/launch':>10s}")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:45]:45s} {v[0]:8d} {v[1]/1e6:10.3f} {100*v[1]/tot:6.1f}% {v[1]/1e6/v[0]:10.4f}")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio","smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]

def rep(path, out_json=None):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")][:60]}
        for w in WANT:
            if w in hdr:
                d[w] = f"{r[hdr.index(w)]} {units[hdr.index(w)]}"
        res.append(d)
        print("---", d["kernel"])
        for w in WANT:
            if w in d: print(f"   {w:85s} {d[w]}")
    if out_json: json.dump(res, open(out_json, "w"), indent=1)

if __name__ == "__main__":
    if sys.argv[1].endswith(".csv"): launches(sys.argv[1])
    else: rep(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
