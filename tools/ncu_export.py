"""Turn an .ncu-rep (one profiled launch) into the tracked summaries under profiles/:
   <out>.csv             metric,unit,value for every metric of `ncu --page raw`
   <out>_stalls.csv      warp-state samples per issue (smsp__average_warps_issue_stalled_*_per_issue_active)

    python tools/ncu_export.py gpurun_out/prof.ncu-rep profiles/r02_name
"""
import csv, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
with open(out + ".csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit", "value"])
    for h, u, v in zip(hdr, units, vals):
        w.writerow([h, u, v])
stalls = []
for h, u, v in zip(hdr, units, vals):
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        try:
            stalls.append((float(v.replace(",", "")), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
tot = sum(x for x, _ in stalls) or 1.0
with open(out + "_stalls.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["warp_state", "warps_per_issue_active", "share_pct"])
    for x, name in sorted(stalls, reverse=True):
        w.writerow([name, f"{x:.4f}", f"{100 * x / tot:.1f}"])
print("wrote", out + ".csv", out + "_stalls.csv")
