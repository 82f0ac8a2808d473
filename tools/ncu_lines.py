"""Per-source-line cost of one kernel of an ncu report (captured with --import-source on, library built with -lineinfo):
   python tools/ncu_lines.py report.ncu-rep <kernel index> [top N]
Prints the lines with the most executed warp instructions and their share of the stall samples."""
import csv, subprocess, sys, io

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", f":::{kid}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
lines = {}
cur_file, hdr = None, None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r; ie = r.index("Instructions Executed"); isamp = r.index("# Samples"); ite = r.index("Thread Instructions Executed"); continue
    if hdr and r[0] not in ("",) and r[2] == "-":
        key = (cur_file, int(r[0]))
        a = lines.setdefault(key, [0.0, 0.0, 0.0, r[1]])
        a[0] += float(r[ie] or 0); a[1] += float(r[isamp] or 0); a[2] += float(r[ite] or 0)
ti = sum(a[0] for a in lines.values()); ts = sum(a[1] for a in lines.values())
print(f"total warp instructions {ti:.0f}, samples {ts:.0f}")
for (f, ln), a in sorted(sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]):
    print(f"{f:14s}:{ln:4d} {100 * a[0] / ti:5.1f}% inst {100 * a[1] / max(ts, 1):5.1f}% samp  thr/inst {a[2] / max(a[0], 1):4.1f}  {a[3].strip()[:110]}")
