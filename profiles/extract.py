#!/usr/bin/env python
"""Turn an .ncu-rep (read with `ncu -i`) into the small text summary committed under profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def main(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"kernel: {name}")
        for k in KEYS:
            if k in hdr:
                print(f"  {k:70s} {r[hdr.index(k)]:>16s} {rows[1][hdr.index(k)]}")
        st = []
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    st.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
                except ValueError:
                    pass
        print("  stall reasons (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
