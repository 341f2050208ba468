"""Print selected metrics from an .ncu-rep (raw page). python tools/ncu_metrics.py rep.ncu-rep [regex ...]"""
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
pats = sys.argv[2:] or [
    r"^gpu__time_duration\.sum$", r"^sm__cycles_elapsed\.max$", r"sm__cycles_elapsed\.max\.per_second",
    r"sm__pipe_tensor_cycles_active_realtime\.avg\.pct", r"^sm__pipe_tensor_cycles_active\.avg\.pct", r"hmma_cycles_active_realtime\.avg$", r"^smsp__issue_active\.avg\.pct_of_peak_sustained_active$",
    r"^smsp__inst_executed\.sum$", r"^sm__inst_executed_pipe_[a-z_]+\.avg\.pct_of_peak_sustained_active$",
    r"^dram__bytes_(read|write)\.sum$", r"^lts__t_bytes\.sum$", r"lts__t_sectors_op_read\.sum$",
    r"^l1tex__data_pipe_lsu_wavefronts\.avg\.pct", r"^smsp__average_warps?_issue_stalled_[a-z_]+_per_issue_active",
    r"^smsp__average_warp_latency_issue_stalled_[a-z_]+\.ratio$", r"launch__registers_per_thread$",
    r"^sm__warps_active\.avg\.pct", r"^gpu__dram_throughput", r"^lts__throughput\.avg\.pct",
    r"^l1tex__throughput\.avg\.pct", r"^sm__throughput\.avg\.pct", r"smsp__warps_eligible\.avg\.per_cycle_active",
    r"^smsp__pcsamp_warps_issue_stalled_[a-z_]+$", r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared",
    r"^sm__mem_tensor", r"smem_throughput|l1tex__data_pipe_tc",
]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u = rows[0], rows[1]
for v in rows[2:]:
    print("##", v[h.index("Kernel Name")][:80] if "Kernel Name" in h else "")
    for i, k in enumerate(h):
        if any(re.search(p, k) for p in pats):
            try:
                if float(v[i].replace(",", "")) == 0:
                    continue
            except ValueError:
                pass
            print(f"  {k:90s} {v[i]:>16s} {u[i]}")
