"""Extract dram__bytes_read.sum + dram__bytes_write.sum (per launch) of every kernel in a set of .ncu-rep files
(`ncu --set full` captures) into profiles/dram_traffic.json -- bench.py reads roofline.traffic from there by kernel
name, so the figure in the bench line is tied to a capture on file, not to a constant in the source.

usage: python tools/ncu_traffic.py gpurun_out/prof_a.ncu-rep [more.ncu-rep ...]"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out_path = os.path.join(ROOT, "profiles", "dram_traffic.json")
try:
    doc = json.load(open(out_path))
except Exception:
    doc = {"how": "ncu --set full --clock-control none, one launch per kernel, 1920x1080 frame (tools/prof_detect.py)", "kernels": {}}
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for vals in rows[2:]:
        name = vals[ix["Kernel Name"]].split("(")[0].replace("void ", "").strip()
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            tot += float(vals[ix[m]]) * UNIT[units[ix[m]]]
        t = float(vals[ix["gpu__time_duration.sum"]]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}[units[ix["gpu__time_duration.sum"]]]
        doc["kernels"][name] = {"dram_bytes": tot, "duration_s_under_ncu": t, "source": os.path.relpath(rep, ROOT)}
json.dump(doc, open(out_path, "w"), indent=1, sort_keys=True)
print(json.dumps(doc, indent=1, sort_keys=True))
