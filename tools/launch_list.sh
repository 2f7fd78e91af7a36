#!/bin/bash
# Per-launch device times of one 1920x1080 detect (cold-cache, serialised: compare shares).  Output: gpurun_out/launches.csv
python tools/prof_detect.py > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${1:-10} -c ${2:-10} --csv --log-file gpurun_out/launches.csv python tools/prof_detect.py > gpurun_out/ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>10]
h=rows[0]; ik=h.index("Kernel Name"); im=h.index("Metric Name"); iv=h.index("Metric Value"); ii=h.index("ID")
d={}
for r in rows[1:]:
    d.setdefault((int(r[ii]), r[ik][:28]),{})[r[im].split(".")[0].replace("gpu__time_duration","us").replace("sm__pipe_fp64_cycles_active","fp64%").replace("dram__bytes_","dram_")]=r[iv]
tot=0
for k,v in sorted(d.items()):
    us=float(v["us"])/1000; tot+=us
    print(f"{k[0]:3d} {k[1]:28s} {us:8.1f} us  fp64 {float(v['fp64%']):5.1f}%  rd {float(v['dram_read'])/1e6:7.1f} MB  wr {float(v['dram_write'])/1e6:7.1f} MB")
print("total", round(tot,1), "us")
PY
