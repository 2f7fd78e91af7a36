"""Shared parity checks between the CUDA engine and the float64 oracle (tests only).

Tolerances are BASELINE.json's north_star:
  * Gaussian / DoG levels within 1e-5 relative of the float64 dense-2D result.  Gaussian levels are
    judged pointwise.  DoG levels are judged pointwise where abs(DoG) >= 0.012 (the smallest value the
    path ever acts on, sift.js:293) and against that same 0.012 floor elsewhere (the
    reference gives no rule, this is the one used).
  * >= 99.5 % of keypoints matched on (octave, scaleLevel, localX, localY) with abs(d absoluteX/Y)
    <= 1e-3 px; every mismatch must come from a candidate within 1e-5 of a threshold or an exact tie.
"""
from __future__ import annotations

import numpy as np

LEVEL_RTOL = 1e-5
DOG_FLOOR = 0.012
POS_TOL = 1e-3


def check_levels(eng, ora, L):
    worst_g = worst_d = 0.0
    n_oct, nlev = eng.pyramid_info()
    assert n_oct == len(ora.gauss) and nlev == len(ora.gauss[0])
    for o in range(n_oct):
        assert eng.octave_size(o) == (ora.gauss[o][0].shape[1], ora.gauss[o][0].shape[0])
        for s in range(nlev):
            g = eng.get_level(L.SIFT_LEVEL_GAUSSIAN, o, s).astype(np.float64)
            ref = ora.gauss[o][s]
            rel = np.abs(g - ref) / np.maximum(np.abs(ref), 1e-3)
            worst_g = max(worst_g, float(rel.max()))
            assert abs(eng.blur_level(L.SIFT_LEVEL_GAUSSIAN, o, s) - ora.blur[o][s]) <= 1e-12 * ora.blur[o][s]
        for s in range(nlev - 1):
            d = eng.get_level(L.SIFT_LEVEL_DOG, o, s).astype(np.float64)
            ref = ora.dog[o][s]
            rel = np.abs(d - ref) / np.maximum(np.abs(ref), DOG_FLOOR)
            worst_d = max(worst_d, float(rel.max()))
    assert worst_g <= LEVEL_RTOL, f"Gaussian level rel err {worst_g:.3g}"
    assert worst_d <= LEVEL_RTOL, f"DoG level rel err {worst_d:.3g}"
    return worst_g, worst_d


def _cand_key(c):
    return (int(c["octave"]), int(c["scale"] if "scale" in c else c["scaleLevel"]), int(c["y"]), int(c["x"]))


def check_candidates(cands, ora, near_tol=1e-5, pix_thr=0.012):
    """cands: structured CANDIDATE_DTYPE in reference order. Mismatches must sit on the pre-filter threshold
    or be float32 ties of the 26-neighbourhood."""
    got = [(int(c["octave"]), int(c["scaleLevel"]), int(c["y"]), int(c["x"])) for c in cands]
    want = [_cand_key(c) for c in ora.candidates]
    assert got == sorted(got), "candidates not in reference order"
    sg, sw = set(got), set(want)
    bad = []
    for k in sg ^ sw:
        o, s, y, x = k
        v = ora.dog[o][s][y, x]
        near_thr = abs(abs(v) - pix_thr) <= near_tol * pix_thr * 10
        nb = np.stack([ora.dog[o][s + d][y - 1:y + 2, x - 1:x + 2] for d in (-1, 0, 1)]).ravel()
        nb = np.delete(nb, 13)
        tie = np.min(np.abs(nb - v)) <= max(abs(v), DOG_FLOOR) * 1e-6   # closer than float32 can tell apart
        if not (near_thr or tie):
            bad.append((k, v))
    assert not bad, f"unexplained candidate mismatches: {bad[:5]}"
    return len(sg & sw), len(sg ^ sw)


def check_keypoints(kps, ora, min_match=0.995):
    want = {}
    for k in ora.keypoints:
        want.setdefault((k["octave"], k["candScale"], k["candY"], k["candX"]), k)
    got = {}
    for k in kps:
        got.setdefault((int(k["octave"]), int(k["candScale"]), int(k["candY"]), int(k["candX"])), k)
    keys_got = [(int(k["octave"]), int(k["candScale"]), int(k["candY"]), int(k["candX"])) for k in kps]
    assert keys_got == sorted(keys_got), "keypoints not in reference (candidate) order"
    matched = 0
    worst = 0.0
    for key, w in want.items():
        g = got.get(key)
        if g is None:
            continue
        same_cell = (int(g["scaleLevel"]) == w["scaleLevel"] and int(g["localX"]) == w["localX"]
                     and int(g["localY"]) == w["localY"])
        dx = abs(float(g["absoluteX"]) - w["absoluteX"])
        dy = abs(float(g["absoluteY"]) - w["absoluteY"])
        if same_cell and dx <= POS_TOL and dy <= POS_TOL:
            matched += 1
            worst = max(worst, dx, dy)
            assert abs(float(g["absoluteSigma"]) - w["absoluteSigma"]) <= 1e-3 * w["absoluteSigma"]
            assert abs(float(g["interpolatedValue"]) - w["interpolatedValue"]) <= 1e-5 * max(abs(w["interpolatedValue"]), DOG_FLOOR)
    total = max(len(want), len(got), 1)
    frac = matched / total
    assert frac >= min_match, f"only {matched}/{total} keypoints matched ({frac:.4f})"
    return matched, total, worst
