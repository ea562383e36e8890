"""Per-source-line hot spots from an ncu report: python tools/ncu_lines.py report.ncu-rep [kernel-regex] [top]
Aggregates SASS rows of `--page source --print-source sass,cuda` by CUDA source line: warp-instructions executed and
stall samples (needs -lineinfo and --import-source on)."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_fn, hdr = None, None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr) or cur_fn is None:
        continue
    if len(sys.argv) > 2 and sys.argv[2] and not re.search(sys.argv[2], cur_fn):
        continue
    d = dict(zip(hdr, r))
    # two columns are named "Source": the first is the CUDA line text, the second the SASS text
    line = r[0]; text = r[1]
    try:
        inst = int(r[hdr.index("Instructions Executed")] or 0)
        samp = int(r[hdr.index("# Samples")] or 0)
    except ValueError:
        continue
    a = agg.setdefault((cur_fn.split("(")[0], line), [text, 0, 0, collections.Counter()])
    a[1] += inst; a[2] += samp
    for k in ("stall_barrier", "stall_long_sb", "stall_short_sb", "stall_mio", "stall_math", "stall_wait", "stall_not_selected", "stall_lg", "stall_branch_resolving", "stall_no_inst", "stall_dispatch"):
        try:
            a[3][k] += int(r[hdr.index(k)] or 0)
        except ValueError:
            pass
tot_i = sum(a[1] for a in agg.values()) or 1
tot_s = sum(a[2] for a in agg.values()) or 1
print(f"total warp-instructions {tot_i}, samples {tot_s}")
for (fn, line), a in sorted(agg.items(), key=lambda x: -x[1][2])[:top]:
    st = ", ".join(f"{k[6:]}={v}" for k, v in a[3].most_common(3) if v)
    print(f"{fn[-24:]:24s} L{line:>4s} inst {a[1] / tot_i * 100:5.1f}%  samp {a[2] / tot_s * 100:5.1f}%  [{st}]  {a[0].strip()[:90]}")
