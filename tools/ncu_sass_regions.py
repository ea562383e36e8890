"""Segments the SASS of one kernel in an ncu report into runs of equal execution frequency (= loops / phases):
python tools/ncu_sass_regions.py report.ncu-rep [min_pct]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows, hdr = [], None
for r in csv.reader(io.StringIO(out)):
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    rows.append((int(r[hdr.index("Instructions Executed")] or 0), int(r[hdr.index("# Samples")] or 0), r[1]))
tot = sum(r[0] for r in rows) or 1
ts = sum(r[1] for r in rows) or 1
print("total warp-instructions", tot, "samples", ts)
i = 0
while i < len(rows):
    j, s, smp, base = i, 0, 0, rows[i][0]
    while j < len(rows) and abs(rows[j][0] - base) <= max(0.0002 * tot, 0.25 * base):
        s += rows[j][0]; smp += rows[j][1]; j += 1
    if s / tot * 100 >= minp:
        ops = {}
        for r in rows[i:j]:
            t = r[2].split()
            op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
            ops[op] = ops.get(op, 0) + 1
        print(f"sass {i:4d}-{j - 1:4d} n={j - i:3d} exec/instr {base / tot * 100:.3f}% total {s / tot * 100:5.1f}% samples {smp / ts * 100:5.1f}% ",
              " ".join(f"{k}:{v}" for k, v in sorted(ops.items(), key=lambda x: -x[1])[:9]))
    i = j
