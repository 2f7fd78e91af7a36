python tools/variant_compare.py "" "SIFT_B200_FUSED0_HI1" 2>&1 | tail -8
echo "== default (hi3)"; python tools/kernel_times.py 2>&1 | tail -14
echo "== HI1"; SIFT_B200_FUSED0_HI1=1 python tools/kernel_times.py 2>&1 | tail -14
echo "== hi3 cap2"; SIFT_B200_OCT0_CTAS=2 python tools/kernel_times.py 2>&1 | tail -14
echo "== hi1 cap2"; SIFT_B200_FUSED0_HI1=1 SIFT_B200_OCT0_CTAS=2 python tools/kernel_times.py 2>&1 | tail -14
