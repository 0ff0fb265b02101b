"""Per-source-line instruction and stall-sample totals of an .ncu-rep (needs -lineinfo and --import-source on).
Usage: ncu_lines.py file.ncu-rep [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
lines = {}
fname = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Name":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        iE, iS = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= max(iE, iS) or not r[0]:
        continue
    try:
        n, s = int(r[iE]), int(r[iS] or 0)
    except ValueError:
        continue
    key = (fname, int(r[0]))
    e = lines.setdefault(key, [0, 0, r[1].strip()])
    e[0] += n; e[1] += s
tot = sum(v[0] for v in lines.values()) or 1
stot = sum(v[1] for v in lines.values()) or 1
print(f"total warp instructions {tot}, samples {stot}")
for (fn, ln), (n, s, src) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{fn}:{ln:5d} {100 * n / tot:5.1f}% inst {100 * s / stot:5.1f}% stall | {src[:110]}")
