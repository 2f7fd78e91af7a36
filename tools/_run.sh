python tools/variant_diff.py "" "SIFT_B200_NO_MMA" 2>&1 | tail -8
echo "== default (mma)"; python tools/kernel_times.py 2>&1 | tail -14
