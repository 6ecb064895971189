"""Summarise an .ncu-rep (raw page) into the handful of counters DESIGN.md / profiles/ quote.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [pattern ...]"""
import csv
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct",
           "gpu__dram_throughput", "sm__pipe_tensor", "sm__inst_executed_pipe_tensor", "sm__throughput.avg.pct",
           "lts__t_bytes.sum", "lts__throughput.avg.pct", "lts__t_sector_hit_rate", "l1tex__throughput.avg.pct",
           "sm__warps_active.avg.pct", "launch__registers_per_thread", "launch__shared_mem", "sm__cycles_elapsed.avg ",
           "sm__cycles_active.avg", "smsp__inst_executed.sum ", "smsp__average_warp", "issue_stalled", "smsp__issue_active",
           "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma", "sm__inst_executed_pipe_lsu", "smsp__inst_issued",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared", "smsp__cycles_active", "launch__grid_size", "clock_rate",
           "lts__t_sectors_srcunit_tex_op_read.sum", "tensor"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== kernel:", r[hdr.index("Kernel Name")], "grid", r[hdr.index("Grid Size")], "block", r[hdr.index("Block Size")])
        for h, u, v in zip(hdr, units, r):
            if any(p.strip() in h for p in pats):
                print(f"{h:100s} {u:14s} {v}")


if __name__ == "__main__":
    main()
