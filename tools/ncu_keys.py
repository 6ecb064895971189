"""Key ncu metrics of a report: python tools/ncu_keys.py <file.ncu-rep>  (runs `ncu -i ... --page raw --csv`)."""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["Kernel Name", "gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in idx:
            print(f"{w}: {r[idx[w]]} {units[idx[w]]}")
    vals = []
    for h in stall:
        try:
            vals.append((float(r[idx[h]]), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        except ValueError:
            pass
    print("top stalls (warps per issue):", [(round(v, 2), n) for v, n in sorted(vals, reverse=True)[:7]])
