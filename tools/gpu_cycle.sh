#!/bin/bash
# One GPU cycle: parity tests, bench, and (optionally) an ncu capture of one kernel.  usage: gpu_cycle.sh [kernel-regex]
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1
python - <<PY
import json
l=[x for x in open("gpurun_out/bench.log") if x.startswith("{")]
if not l: print(open("gpurun_out/bench.log").read()[-2000:])
else:
    d=json.loads(l[-1]); print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "ms/frame", round(d["ms_per_step"]/d["config"]["frames_per_gpu_per_step"],4))
    print({k:round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()})
    print("frac dom", round(d["roofline"]["frac"],3), "whole", round(d["roofline"]["whole_path"]["frac"],3), d["clocks"], "launches", d["gpu_launches"])
PY
if [ -n "$1" ]; then
  python tools/prof_detect.py > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:$1 -s ${2:-1} -c 1 -o gpurun_out/prof_$1 python tools/prof_detect.py > gpurun_out/ncu.log 2>&1
  tail -n 2 gpurun_out/ncu.log
fi
