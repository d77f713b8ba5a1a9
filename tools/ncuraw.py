"""Key metrics + stall reasons of the first kernel in an ncu report (raw page)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for vals in rows[2:]:
    print("kernel:", vals[hdr.index("Kernel Name")][:90])
    want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
            'l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.sum', 'sm__cycles_active.avg',
            'l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_active',
            'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
            'smsp__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active']
    for w in want:
        if w in hdr: print(f"  {w:78s} {vals[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    st = []
    for i, h in enumerate(hdr):
        if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h:
            try: st.append((float(vals[i]), h.split('stalled_')[-1].replace('_per_issue_active.ratio', '')))
            except ValueError: pass
    print("  stalls per issue:", ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)[:9]))
