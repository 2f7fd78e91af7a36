#!/bin/bash
# Rebuild ONE object with different -D knobs on the GPU box and bench each variant.
# usage: variant_bench.sh <object-stem> "<EXTRA flags 1>" "<EXTRA flags 2>" ...      (run under gpurun)
stem=$1; shift
C=sift-scale-space-extrema-detection_b200/csrc
cp $C/../libsift_b200.so /tmp/lib_orig.so
for extra in "$@"; do
  for st in ${stem//,/ }; do rm -f $C/$st.o; done
  if ! make -C $C EXTRA="$extra" > gpurun_out/variant_build.log 2>&1; then echo "BUILD FAILED: $extra"; tail -5 gpurun_out/variant_build.log; continue; fi
  grep -A2 "scan_tma_kernelILi5ELb0" gpurun_out/variant_build.log | grep -o "Used [0-9]* registers.*" | head -1
  python bench.py --headline-only --steps 5 --warmup 3 > gpurun_out/bench_variant.log 2>&1
  python - "$extra" <<'PY'
import json, sys
l=[x for x in open("gpurun_out/bench_variant.log") if x.startswith("{")]
if not l: print("FAILED", sys.argv[1], open("gpurun_out/bench_variant.log").read()[-600:])
else:
    d=json.loads(l[-1]); print(repr(sys.argv[1]), "value", round(d["value"],1), {k:round(v["ms_per_frame"],4) for k,v in d["roofline"]["kernels"].items()})
PY
done
for st in ${stem//,/ }; do rm -f $C/$st.o; done; cp /tmp/lib_orig.so $C/../libsift_b200.so
