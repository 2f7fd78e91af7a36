#!/bin/bash
# Per-kernel SASS evidence for the Blackwell-specific paths: TMA loads / stores (UTMALDG / UTMASTG), mbarrier
# transactions (SYNCS), async-proxy fences, cp.async (LDGSTS), the fp64 FMA count and the fp64 matrix instruction (DMMA.8x8x4).  Static counts from the built
# library (cuobjdump -sass), written to profiles/r02_sass_counts.txt.
lib=${1:-sift-scale-space-extrema-detection_b200/libsift_b200.so}
cuobjdump -sass "$lib" | awk '
  /Function :/ { fn=$3; next }
  /^[ \t]+\/\*[0-9a-f]+\*\// {
    total[fn]++
    if ($0 ~ /UTMALDG/) ldg[fn]++
    if ($0 ~ /UTMASTG/) stg[fn]++
    if ($0 ~ /SYNCS/) syncs[fn]++
    if ($0 ~ /FENCE.VIEW.ASYNC/) fence[fn]++
    if ($0 ~ /LDGSTS/) ldgsts[fn]++
    if ($0 ~ /DFMA/) dfma[fn]++
    if ($0 ~ /DMMA/) dmma[fn]++
    if ($0 ~ /F2F.F32.F64/) f2f[fn]++
  }
  END {
    printf "%-64s %7s %7s %7s %7s %6s %6s %6s %6s %6s\n", "kernel", "instrs", "DFMA", "DMMA", "F2F", "UTMALDG", "UTMASTG", "SYNCS", "FENCE", "LDGSTS"
    for (f in total) printf "%-64s %7d %7d %7d %7d %6d %6d %6d %6d %6d\n", substr(f,1,64), total[f], dfma[f], dmma[f], f2f[f], ldg[f], stg[f], syncs[f], fence[f], ldgsts[f]
  }' | (read hdr; echo "$hdr"; sort)
