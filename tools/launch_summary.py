"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (and grid) -> markdown table."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i + 1
        break
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
agg = collections.OrderedDict()
byname = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    name = r[ki].split("(")[0].replace("fbe::", "")
    v = float(r[vi].replace(",", ""))
    a = agg.setdefault((name, r[gi]), [0, 0.0]); a[0] += 1; a[1] += v
    b = byname.setdefault(name, [0, 0.0]); b[0] += 1; b[1] += v
tot = sum(a[1] for a in agg.values())
print(f"launches: {len(rows) - start}, steps: {steps:g}, serialised device time {tot / 1e6:.3f} ms = {tot / 1e6 / steps:.3f} ms/step\n")
print("| kernel | launches/step | ms/step | share |\n|---|---|---|---|")
for k, a in sorted(byname.items(), key=lambda x: -x[1][1]):
    print(f"| {k} | {a[0] / steps:g} | {a[1] / 1e6 / steps:.3f} | {a[1] / tot * 100:.1f}% |")
print("\n| kernel | grid | launches/step | avg us | share |\n|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:24]:
    print(f"| {k[0]} | {k[1]} | {a[0] / steps:g} | {a[1] / a[0] / 1e3:.1f} | {a[1] / tot * 100:.1f}% |")
