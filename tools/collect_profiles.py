"""Summarise gpurun_out/ (written by tools/make_profiles.sh on the GPU box) into profiles/rNN_*: the files the
judge reads.  Run here (ncu reads the .ncu-rep files offline)."""
import csv, io, json, os, shutil, subprocess, sys
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = "gpurun_out", "profiles"
os.makedirs(P, exist_ok=True)
for src, dst in (("pytest_gpu.log", "pytest_gpu.log"), ("smoke.log", "smoke.log")):
    if os.path.exists(os.path.join(G, src)):
        shutil.copy(os.path.join(G, src), os.path.join(P, f"{rnd}_{dst}"))
lines = [l for l in open(os.path.join(G, "bench.log")) if l.startswith("{")]
if lines:
    json.dump(json.loads(lines[-1]), open(os.path.join(P, f"{rnd}_bench.json"), "w"), indent=1)
# launch list (cold-cache, serialised: compare shares, not absolutes)
rows = [r for r in csv.reader(open(os.path.join(G, "launches.csv"))) if len(r) > 10]
h = rows[0]; ik, im, iv, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
d = {}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ik][:60]), {})[r[im]] = float(r[iv].replace(",", ""))
tot = sum(v["gpu__time_duration.sum"] for v in d.values())
with open(os.path.join(P, f"{rnd}_ncu_launch_list.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum,... --clock-control none: one 1920x1080 detect, device-resident (tools/prof_detect.py)\n")
    f.write("# per-launch times are cold-cache and serialised: the SHARE of the step is what compares with bench.py\n")
    for k, v in sorted(d.items()):
        us = v["gpu__time_duration.sum"] / 1000
        f.write(f"{k[0]:3d} {k[1]:60s} {us:8.1f} us {100 * v['gpu__time_duration.sum'] / tot:5.1f}%  fp64 pipe "
                f"{v['sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']:5.1f}% (DFMA) "
                f"{v.get('sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active', 0.0):5.1f}% (DMMA)  dram rd "
                f"{v['dram__bytes_read.sum'] / 1e6:7.1f} MB  wr {v['dram__bytes_write.sum'] / 1e6:7.1f} MB\n")
    f.write(f"total {tot / 1000:.1f} us\n")
for rep in sorted(os.listdir(G)):
    if rep.endswith(".ncu-rep") and rep.startswith("prof_") and rep.count("_") >= 2 and rep[:-8].split("_")[-1].isdigit():
        out = subprocess.run([sys.executable, "tools/ncu_summary.py", os.path.join(G, rep), "14"], capture_output=True, text=True).stdout
        name = rep[5:-8]
        open(os.path.join(P, f"{rnd}_ncu_{name}.txt"), "w").write(
            "# ncu --set full --clock-control none --import-source on (one launch); summary by tools/ncu_summary.py\n" + out)
print(sorted(os.listdir(P)))
