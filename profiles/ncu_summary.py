#!/usr/bin/env python
"""Summarise an ncu report: `python profiles/ncu_summary.py gpurun_out/x.ncu-rep [pattern ...]`
prints, per profiled launch, the metrics whose names contain one of the patterns (a default set if none)."""
import csv
import io
import subprocess
import sys

DEFAULT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
           "sm__warps_active.avg.pct_of_peak", "launch__registers_per_thread", "launch__occupancy_limit",
           "sm__throughput.avg.pct", "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "inst_executed_pipe_alu.sum",
           "inst_executed_pipe_fma.sum", "inst_executed_pipe_lsu.sum", "inst_executed_pipe_xu.sum", "shared_ld.sum", "bank_conflicts",
           "issue_stalled", "thread_inst_executed_per_inst", "sm__cycles_elapsed.avg ", "launch__shared_mem_per_block", "lts__t_sector_hit_rate"]


def main():
    rep = sys.argv[1]
    pats = sys.argv[2:] or DEFAULT
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print(f"=== {r[4][:100]}  grid {r[8]} block {r[7]}")
        for i, h in enumerate(hdr):
            if any(p in h for p in pats) and r[i] not in ("", "0", "n/a"):
                if "issue_stalled" in h and ("not_issued" in h or float(r[i].replace(",", "")) < 3):
                    continue
                print(f"  {h:90s} {r[i]:>16s} {units[i]}")


if __name__ == "__main__":
    main()
