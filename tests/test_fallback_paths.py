"""The kernels that exist for generality (generic separable blur, cp.async pass B, register-resident
octave-0 tile, non-TMA scan, 8-output FIR shape) are normally reached only by unusual sizes.  Each one can be forced
with an environment knob that the library reads once per process, so every variant runs in its own interpreter and
is held to the same parity bars against the float64 oracle as the default path."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import sys
sys.path.insert(0, %(root)r); sys.path.insert(0, %(root)r + "/tests")
import numpy as np
import oracle, sift_b200
from sift_b200 import _lib as L, fixtures
from parity import check_candidates, check_keypoints, check_levels
eng = sift_b200.Engine(0)
for (w, h, n_oct, seed) in ((208, 144, 3, 3), (1200, 900, 2, 8)):     # the second one is > 1 Mpixel in octave 1: big FIR shape
    u8 = fixtures.synthetic_u8(w, h, seed)
    ora = oracle.detect(fixtures.to_float(u8), numberOfOctaves=n_oct, minBlurLevel=1.6, separable=True, keep_levels=(w < 400))
    prm = L.default_params(numberOfOctaves=n_oct, minBlurLevel=1.6)
    if w < 400:
        eng.build_scale_space(u8, prm)
        check_levels(eng, ora, L)
        cands, _ = eng.find_candidates()
        check_candidates(cands, ora)
    kps, st = eng.detect(u8, prm)
    matched, total, worst = check_keypoints(kps, ora)
    print("OK", w, h, matched, total, st["kernelLaunches"])
"""

# SIFT_B200_OCT0_WS: the persistent warp-specialised octave-0 kernel of blur_oct0.cu (TMA source boxes, TMA stores);
# "SIFT_B200_OCT0_WS+SIFT_B200_NO_TMA_BLUR": the same kernel with the source tile loaded by plain loads
KNOBS = ["", "SIFT_B200_FORCE_GENERIC", "SIFT_B200_FORCE_OLD", "SIFT_B200_NO_TMA", "SIFT_B200_NO_TMA_BLUR",
         "SIFT_B200_FUSED0_LO", "SIFT_B200_FIR_NO8", "SIFT_B200_OCT0_WS", "SIFT_B200_OCT0_WS+SIFT_B200_NO_TMA_BLUR",
         "SIFT_B200_OCT0_SMALL", "SIFT_B200_OCT0_BANDS", "SIFT_B200_OCT0_BANDS+SIFT_B200_OCT0_BAND=64",
         # round 2, DMMA kernels are the default: the scalar-FMA kernels everywhere / the two scalar octave-0 tile
         # kernels / scalar passes for the small octaves / DMMA kernels with copied instead of TMA-loaded tiles
         "SIFT_B200_NO_MMA", "SIFT_B200_FUSED0_HI1", "SIFT_B200_FUSED0_HI3", "SIFT_B200_MMA_BIG_ONLY", "SIFT_B200_MMA_NO_TMA"]


@pytest.mark.gpu
@pytest.mark.parametrize("knob", KNOBS)
def test_forced_kernel_variant_meets_the_parity_bars(knob):
    env = dict(os.environ)
    for k in KNOBS:
        for part in k.split("+"):
            env.pop(part.partition("=")[0], None)
    for part in knob.split("+"):
        if part:
            name, _, val = part.partition("=")
            env[name] = val or "1"
    r = subprocess.run([sys.executable, "-c", SCRIPT % {"root": ROOT}], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (knob, r.stdout[-2000:], r.stderr[-3000:])
    assert r.stdout.count("OK ") == 2
