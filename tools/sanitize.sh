#!/bin/bash
# usage: tools/sanitize.sh memcheck|racecheck   (one tool per gpurun call: /opt/skills/guides/B200_PROFILING.md)
# Runs tools/sanitize_cases.py under compute-sanitizer for the default kernels and for the forced variants;
# the summaries land in gpurun_out/sanitizer_<tool>.txt (copy to profiles/).
tool=${1:-memcheck}
out=gpurun_out/sanitizer_${tool}.txt
: > $out
for knob in "" SIFT_B200_OCT0_WS SIFT_B200_FUSED0_LO SIFT_B200_NO_TMA_BLUR SIFT_B200_FORCE_GENERIC; do
  echo "=== compute-sanitizer --tool $tool, variant: ${knob:-default}" >> $out
  if [ -n "$knob" ]; then export $knob=1; fi
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py 2>&1 | grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize cases done|Error|error|Race|hazard" | head -40 >> $out
  if [ -n "$knob" ]; then unset $knob; fi
done
cat $out
