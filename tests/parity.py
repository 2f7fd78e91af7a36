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
NEAR_TOL = 1e-5      # north_star: a mismatch must lie within 1e-5 (relative) of a threshold, or be an exact tie


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


def _explain_candidate(key, ora, pix_thr, near_tol):
    """Why may an extremum be on one list and not the other?  Only because abs(value) sits on the pre-filter
    threshold (sift.js:293) or because two of its 27 voxels are closer than float32 storage can tell apart
    (sift.js:261,266 ties).  Returns the reason or None."""
    o, s, y, x = key
    v = ora.dog[o][s][y, x]
    if abs(abs(v) - pix_thr) <= near_tol * pix_thr:
        return "threshold"
    nb = np.stack([ora.dog[o][s + d][y - 1:y + 2, x - 1:x + 2] for d in (-1, 0, 1)]).ravel()
    nb = np.delete(nb, 13)
    if np.min(np.abs(nb - v)) <= max(abs(v), DOG_FLOOR) * 1e-6:       # closer than float32 can tell apart
        return "tie"
    return None


def check_candidates(cands, ora, near_tol=NEAR_TOL, pix_thr=0.012, low=False):
    """cands: structured CANDIDATE_DTYPE in reference order, compared with the oracle's candidate list (or, with
    low=True, its low-contrast list).  Every mismatch must be explained by _explain_candidate."""
    got = [(int(c["octave"]), int(c["scaleLevel"]), int(c["y"]), int(c["x"])) for c in cands]
    want = [_cand_key(c) for c in (ora.low_contrast if low else ora.candidates)]
    assert got == sorted(got), "candidates not in reference order"
    sg, sw = set(got), set(want)
    assert len(sg) == len(got), "duplicate candidates"
    bad = [(k, float(ora.dog[k[0]][k[1]][k[2], k[3]])) for k in sg ^ sw
           if _explain_candidate(k, ora, pix_thr, near_tol) is None]
    assert not bad, f"unexplained candidate mismatches: {bad[:5]}"
    # values are the float32 DoG samples themselves
    by_key = {_cand_key(c): c["value"] for c in (ora.low_contrast if low else ora.candidates)}
    for c, k in zip(cands, got):
        if k in by_key:
            assert abs(float(c["value"]) - by_key[k]) <= LEVEL_RTOL * max(abs(by_key[k]), DOG_FLOOR)
    return len(sg & sw), len(sg ^ sw)


def _kp_key(k):
    return (int(k["octave"]), int(k["candScale"]), int(k["candY"]), int(k["candX"]))


def check_keypoints(kps, ora, min_match=0.995, near_tol=NEAR_TOL, pix_thr=0.012, device_dog=None, params=None):
    """>= 99.5 % of the keypoints on the same cell within 1e-3 px -- and EVERY record that is not matched must be
    explained, by one of:
      threshold  its candidate sits on the pre-filter threshold / a tie (so the candidate lists differ), or the
                 oracle's refinement walk of that candidate came within near_tol (1e-5) of flipping a decision
                 (oracle_refine_margin: offset bound 0.6, contrast 0.015, edge 12.1, Math.round boundary);
      storage    the quadratic fit at that sample is so ill-conditioned that rounding the DoG levels to the float32
                 the device stores (6e-8 relative, 170x inside the 1e-5 level tolerance) moves the walk: shown by
                 running the ORACLE's float64 refinement (background.js:455-685) on the device's own stored DoG
                 levels (`device_dog(octave) -> levels`) and getting exactly the device's record / rejection.
    Returns (matched, total, worst position error); the explanation counts are attached as check_keypoints.last."""
    want = {}
    for k in ora.keypoints:
        want.setdefault((k["octave"], k["candScale"], k["candY"], k["candX"]), k)
    got = {}
    for k in kps:
        got.setdefault(_kp_key(k), k)
    keys_got = [_kp_key(k) for k in kps]
    assert keys_got == sorted(keys_got), "keypoints not in reference (candidate) order"
    assert len(want) == len(ora.keypoints) and len(got) == len(kps), "two records from one candidate"
    matched = 0
    worst = 0.0
    unmatched = []
    for key, w in want.items():
        g = got.get(key)
        if g is None:
            unmatched.append(key)
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
        else:
            unmatched.append(key)
    unmatched += [k for k in got if k not in want]
    unexplained = []
    reasons = {"threshold": 0, "storage": 0}
    dog_cache = {}
    import oracle as _oracle
    for key in unmatched:
        o, s, y, x = key
        margin = ora.margins.get(key)
        if margin is None:
            # the oracle never refined this candidate: the candidate lists differ -> must be a threshold / tie case
            if ora.dog and _explain_candidate(key, ora, pix_thr, near_tol) is None:
                unexplained.append((key, "candidate only on the device"))
            else:
                reasons["threshold"] += 1
        elif margin <= near_tol:
            reasons["threshold"] += 1
        else:
            why = f"walk margin {margin:.3g}"
            if device_dog is not None:
                if o not in dog_cache:
                    dog_cache[o] = [np.asarray(d, dtype=np.float64) for d in device_dog(o)]
                g = got.get(key)
                cand = {"scale": s, "x": x, "y": y, "value": float(dog_cache[o][s][y, x])}
                outcome, rec = _oracle.refine_on_dog(dog_cache[o], o, cand, params)
                if g is None:
                    same = outcome != "accepted"
                else:
                    same = (rec is not None and rec["scaleLevel"] == int(g["scaleLevel"]) and rec["localX"] == int(g["localX"])
                            and rec["localY"] == int(g["localY"])
                            and abs(rec["absoluteX"] - float(g["absoluteX"])) <= 1e-9 * max(1.0, abs(rec["absoluteX"]))
                            and abs(rec["absoluteY"] - float(g["absoluteY"])) <= 1e-9 * max(1.0, abs(rec["absoluteY"])))
                if same:
                    reasons["storage"] += 1
                    continue
                why += f"; oracle refinement of the device's stored levels gives {outcome}, the device {'a record' if g is not None else 'none'}"
            unexplained.append((key, why))
    check_keypoints.last = dict(reasons, unmatched=len(unmatched))
    assert not unexplained, f"unexplained keypoint mismatches: {unexplained[:5]}"
    total = max(len(want), len(got), 1)
    frac = matched / total
    assert frac >= min_match, f"only {matched}/{total} keypoints matched ({frac:.4f})"
    return matched, total, worst
