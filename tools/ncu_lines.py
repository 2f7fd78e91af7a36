"""Per CUDA source line instruction counts of an .ncu-rep: python tools/ncu_lines.py rep [min-share%]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; lines = []; cur_file = ""
for r in rows:
    if r and r[0] == "File Path": cur_file = r[1].split("/")[-1]
    elif r and r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; ie = hdr.index("Instructions Executed"); ismp = hdr.index("# Samples")
    elif hdr and len(r) >= len(hdr) and r[0] not in ("", "Line No"):
        try: lines.append((cur_file, int(r[0]), r[1].strip()[:100], int(float(r[ie] or 0)), int(float(r[ismp] or 0))))
        except ValueError: pass
tot = sum(l[3] for l in lines) or 1
print("total warp instructions attributed to source lines:", tot)
for f, ln, src, n, smp in lines:
    if 100.0 * n / tot >= thr: print(f"{100.0*n/tot:6.2f}% {smp:6d}  {f}:{ln:<4d} {src}")
