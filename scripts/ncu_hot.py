"""Top SASS instructions by warp-stall samples from an .ncu-rep (source page), with the dominant stall reason."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
kern = None; hdr = None; data = []
for r in rows:
    if r and r[0] == "Kernel Name": kern = r[1]; continue
    if r and r[0] == "Address": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    d = dict(zip(hdr, r)); d["kernel"] = kern; data.append(d)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for k in sorted(set(d["kernel"] for d in data)):
    dd = [d for d in data if d["kernel"] == k]
    tot = sum(int(d["# Samples"]) for d in dd)
    print(f"== {k[:80]}  total samples {tot}, instructions executed {sum(int(d['Instructions Executed']) for d in dd)}")
    for i, d in enumerate(dd): d["idx"] = i
    for d in sorted(dd, key=lambda d: -int(d["# Samples"]))[:top]:
        n = int(d["# Samples"])
        why = max(stalls, key=lambda s: int(d[s] or 0))
        print(f"{d['idx']:5d} {n:7d} {100*n/max(tot,1):5.1f}%  exec={int(d['Instructions Executed']):9d}  {why:18s} {d['Source'].strip()[:90]}")
