"""Per-instruction execution shares of the kernels in an .ncu-rep (source page): python tools/ncu_regions.py rep [kernel-index] [min-share%]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; which = int(sys.argv[2]) if len(sys.argv) > 2 else 0; thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
hi = starts[which]; end = starts[which + 1] - 1 if which + 1 < len(starts) else len(rows)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
lines = []
for r in rows[hi + 1:end]:
    if len(r) < len(hdr): continue
    try: n = int(float(r[ix["Instructions Executed"]] or 0))
    except ValueError: continue
    lines.append((n, r[ix["Source"]][:80], int(float(r[ix["# Samples"]] or 0))))
tot = sum(l[0] for l in lines) or 1
print("total warp instructions", tot)
skipped = 0
for i, (n, s, smp) in enumerate(lines):
    if n / tot * 100 >= thr:
        if skipped: print(f"      ... {skipped} lines below {thr}%"); skipped = 0
        print(f"{i:5d} {n/tot*100:5.2f}% {smp:5d} {s}")
    elif n: skipped += 1
