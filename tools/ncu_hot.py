"""Top stall sites of an ncu source-page CSV: python tools/ncu_hot.py rep.ncu-rep [N]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = his[0]
h = rows[hi]
col = {k: i for i, k in enumerate(h)}
end = next((i for i in range(hi + 1, len(rows)) if not rows[i] or rows[i][0] == "Kernel Name"), len(rows))  # first launch only
data = [r for r in rows[hi + 1:end] if len(r) == len(h)]
stalls = ["stall_barrier", "stall_branch_resolving", "stall_long_sb", "stall_math", "stall_membar", "stall_mio",
          "stall_not_selected", "stall_selected", "stall_short_sb", "stall_wait", "stall_dispatch", "stall_no_inst",
          "stall_lg", "stall_sleep"]
tot = sum(int(r[col["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {s: sum(int(r[col[s]] or 0) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
data.sort(key=lambda r: -int(r[col["# Samples"]] or 0))
for r in data[:n]:
    s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[k]] or 0), k) for k in stalls), reverse=True)[:2]
    print(f"{s:7d} {100.0 * s / tot:5.1f}%  {r[col['Source']][:70]:70s} {top[0][1][6:]}={top[0][0]} {top[1][1][6:]}={top[1][0]}")
