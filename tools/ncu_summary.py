"""One-line-per-kernel summary of an ncu report: python tools/ncu_summary.py report.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(out)))
h = r[0]
want = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conflicts"),
        ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"), ("launch__registers_per_thread", "regs"),
        ("sm__inst_executed.sum", "warp_inst"), ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "gld_sectors"), ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "gld_requests")]
for row in r[2:]:
    parts = []
    for k, name in want:
        if k in h:
            v = row[h.index(k)]
            u = r[1][h.index(k)]
            if name == "kernel":
                v = v.split("(")[0].replace("fbe::", "")
            parts.append(f"{name}={v}{(' ' + u) if u and name in ('us', 'dram_rd', 'dram_wr') else ''}")
    print("  ".join(parts))
