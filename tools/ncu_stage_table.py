"""Markdown table, one row per captured kernel launch, from `ncu -i rep --page raw --csv` output:
   python tools/ncu_stage_table.py stage_raw.csv"""
import csv, sys
r = list(csv.reader(open(sys.argv[1])))
h, u = r[0], r[1]
cols = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "DRAM rd"), ("dram__bytes_write.sum", "DRAM wr"),
        ("dram__bytes.sum.per_second", "DRAM rate"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"), ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/smem pipe %"),
        ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue %"), ("sm__inst_executed.sum.per_cycle_elapsed", "IPC (GPU)"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"), ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"), ("launch__registers_per_thread", "regs")]
cols = [(k, n) for k, n in cols if k in h]
print("| kernel | grid | " + " | ".join(n for _, n in cols) + " |")
print("|---|---|" + "---|" * len(cols))
ki, gi = h.index("Kernel Name"), h.index("Grid Size")
for row in r[2:]:
    out = []
    for k, n in cols:
        i = h.index(k)
        v = row[i].replace(",", "")
        try:
            f = float(v)
            v = f"{f:.3g}" if f < 1000 else f"{f:.0f}"
        except ValueError:
            pass
        unit = u[i]
        out.append(v + (" " + unit if unit and unit not in ("%", "inst/cycle", "register/thread") else ""))
    print(f"| {row[ki].split('(')[0].replace('fbe::', '')} | {row[gi]} | " + " | ".join(out) + " |")
