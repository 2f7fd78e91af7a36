"""Small detections that reach every kernel variant, for compute-sanitizer (one tool per run):
   compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
   compute-sanitizer --tool racecheck python tools/sanitize_cases.py
Sizes: the smoke case, odd sizes (97x61: ceil halving, 5x37: image narrower than every kernel), scalesPerOctave = 5
(non-TMA scan), a strip pair with the C-ABI halo exchange, the stage API with the low-contrast list, the ordered
device-resident call.  SIFT_B200_* knobs in the environment select the kernel variant (tools/sanitize.sh loops them)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sift_b200
from sift_b200 import _lib as L, fixtures, mosaic

eng = sift_b200.Engine(0)
n = 0
for (w, h, n_oct, spo) in ((160, 120, 3, 3), (97, 61, 3, 3), (5, 37, 2, 3), (64, 48, 2, 5), (200, 150, 4, 3)):
    u8 = fixtures.synthetic_u8(w, h, 7 + w)
    prm = L.default_params(numberOfOctaves=n_oct, scalesPerOctave=spo, minBlurLevel=1.6)
    kps, st = eng.detect(u8, prm)
    eng.build_scale_space(u8.astype(np.float32) / 255, prm)
    c, low = eng.find_candidates(want_low_contrast=True)
    k2, _ = eng.refine(c, prm)
    n += len(kps) + len(k2)
frames = np.stack([fixtures.synthetic_u8(96, 80, 50 + i) for i in range(5)])
prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
kb, offs, _ = eng.detect_batch(frames, prm)
n += len(kb)
u8 = fixtures.synthetic_u8(64, 400, 3)
engines = [sift_b200.Engine(0) for _ in range(2)]
km, _, _ = mosaic.detect_mosaic_local(engines, u8, prm, margin=8)
n += len(km)
for e in engines:
    e.close()
eng.close()
print("sanitize cases done:", n, "records")
