"""Turn ncu outputs into the markdown summaries under profiles/ (generated, not typed):
   python tools/ncu_summaries.py launches <launches.csv> <steps in the capture> > profiles/rX_launches.md
   python tools/ncu_summaries.py stage <stage.ncu-rep> > profiles/rX_stage_ncu.md"""
import collections, csv, subprocess, sys


def launches(path, nsteps):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    per, grids = collections.OrderedDict(), collections.OrderedDict()
    n = 0
    for r in rows[hdr + 2:]:
        if len(r) < 15:
            continue
        name = r[4].split("(")[0].replace("fbe::", "").replace("void ", "")
        ns = float(r[-1]); n += 1
        per.setdefault(name, []).append(ns)
        grids.setdefault((name, r[8]), []).append(ns)
    tot = sum(sum(v) for v in per.values())
    print(f"launches: {n}, steps: {nsteps}, serialised device time {tot / 1e6:.3f} ms = {tot / 1e6 / nsteps:.3f} ms/step\n")
    print("| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        print(f"| {k} | {len(v) / nsteps:.2g} | {sum(v) / 1e6 / nsteps:.3f} | {100 * sum(v) / tot:.1f}% |")
    print("\n| kernel | grid | launches/step | avg us | share |\n|---|---|---|---|---|")
    for (k, g), v in sorted(grids.items(), key=lambda kv: -sum(kv[1]))[:26]:
        print(f"| {k} | {g} | {len(v) / nsteps:.2g} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")


def stage(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "time us"), ("smsp__inst_executed.sum", "warp-instr M"),
            ("dram__bytes_read.sum", "DRAM rd MB"), ("dram__bytes_write.sum", "DRAM wr MB"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU %"),
            ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA %"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU %"),
            ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU %"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
            ("launch__registers_per_thread", "regs")]
    units = dict(zip(hdr, rows[1]))
    print("| " + " | ".join(c[1] for c in cols) + " | DRAM TB/s |\n|" + "---|" * (len(cols) + 1))
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        vals = []
        for key, label in cols:
            v = d.get(key, "")
            if key == "Kernel Name":
                v = v.split("(")[0].replace("fbe::", "").replace("void ", "")
            elif key == "gpu__time_duration.sum":
                t = float(v) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "msecond": 1e3, "usecond": 1.0, "nsecond": 1e-3}.get(units[key], 1.0)
                v = f"{t:.1f}"
            elif key == "smsp__inst_executed.sum":
                v = f"{float(v) / 1e6:.1f}"
            elif key.startswith("dram__bytes"):
                f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(units[key], 1.0)
                v = f"{float(v) * f:.1f}"
            elif "pct" in key:
                v = f"{float(v):.1f}"
            vals.append(v)
        tus = float(vals[2])
        bw = (float(vals[4]) + float(vals[5])) / tus if tus > 0 else 0.0      # MB / us = TB/s
        print("| " + " | ".join(vals) + f" | {bw:.2f} |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], float(sys.argv[3]))
    else:
        stage(sys.argv[2])
