"""GPU parity at BASELINE.json's own sizes, against the separable float64 oracle.

tests/test_oracle_kat.py pins the separable oracle to the dense 2D one (the reference's algorithm, sift.js:96-139) at
~1e-14 on small images; here it stands in for it on 1920x1080 / 4 octaves (configs[1], the headline), one 1280x720
frame (configs[2]), 3840x2160 / 6 octaves (configs[3]: radii 231 ... 463 in octaves 4-5, background.js:156-177) and a
mosaic cut into strips (configs[4]), checking all three products -- levels, candidate lists (incl. the low-contrast
list, sift.js:301-306) and refined keypoints (background.js:455-685) -- with every mismatch explained
(tests/parity.py).
"""
import numpy as np
import pytest

import oracle
import sift_b200
from sift_b200 import _lib as L, fixtures
from parity import check_candidates, check_keypoints, check_levels

pytestmark = pytest.mark.gpu


def _full_check(engine, w, h, n_oct, seed, levels=True):
    u8 = fixtures.synthetic_u8(w, h, seed)
    ora = oracle.detect(fixtures.to_float(u8), numberOfOctaves=n_oct, minBlurLevel=1.6, separable=True)
    try:
        assert ora.outcomes["singular"] == 0, "fixture would crash the reference (SURVEY.md Q7)"
        prm = L.default_params(numberOfOctaves=n_oct, minBlurLevel=1.6)
        engine.build_scale_space(u8, prm)
        worst = check_levels(engine, ora, L) if levels else None
        cands, low = engine.find_candidates(want_low_contrast=True)
        same, diff = check_candidates(cands, ora)
        same_low, diff_low = check_candidates(low, ora, low=True)
        assert same + diff >= len(ora.candidates) and same_low + diff_low >= ora.n_low_contrast
        kps, stats = engine.detect(u8, prm)
        # (sift_detect leaves its pyramid in the context: the stored DoG levels can be read back for explanations)
        oprm = oracle.default_params(numberOfOctaves=n_oct, minBlurLevel=1.6)
        matched, total, worst_pos = check_keypoints(
            kps, ora, params=oprm,
            device_dog=lambda o: [engine.get_level(L.SIFT_LEVEL_DOG, o, s) for s in range(5)])
        assert stats["rejSingular"] == 0
        r = {"levels": worst, "cands": (same, diff), "low": (same_low, diff_low), "kps": (matched, total, worst_pos),
             "explained": dict(check_keypoints.last)}
        print(f"{w}x{h}/{n_oct} oct seed {seed}: {r}")
        return r
    finally:
        ora.close()


@pytest.mark.parametrize("seed", [1234, 1235])
def test_1080p_4_octaves(engine, seed):
    """BASELINE configs[1]: the headline frame.  Radii 5 ... 116."""
    r = _full_check(engine, 1920, 1080, 4, seed)
    assert r["kps"][1] > 5000


def test_720p_frame(engine):
    """BASELINE configs[2]: one frame of the batch."""
    r = _full_check(engine, 1280, 720, 4, 1234 + 17)
    assert r["kps"][1] > 2000


def test_2160p_6_octaves(engine):
    """BASELINE configs[3]: octaves 4 and 5 run the radius-generic kernels (R = 59 ... 463 on 480x270 / 240x135)."""
    r = _full_check(engine, 3840, 2160, 6, 1234)
    assert r["kps"][1] > 20000


def test_low_contrast_list_small(engine):
    """the pyramid scan's second compaction output equals SIFT_findExtremas' lowContrastKeypoints
    (sift.js:301-306) over the scales background.js:374-377 visits -- list, order and count."""
    u8 = fixtures.synthetic_u8(320, 240, 77)
    ora = oracle.detect(fixtures.to_float(u8), numberOfOctaves=4, minBlurLevel=1.6, separable=True)
    prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    engine.build_scale_space(u8, prm)
    cands, low = engine.find_candidates(want_low_contrast=True)
    same, diff = check_candidates(low, ora, low=True)
    assert ora.n_low_contrast > 50 and same >= ora.n_low_contrast - diff
    # counting without materialising the list gives the same number
    import ctypes as C
    n, nl = C.c_int(), C.c_int()
    out = np.zeros(len(cands) + 8, dtype=L.CANDIDATE_DTYPE)
    rc = engine._lib.sift_find_candidates(engine._h, None, out.ctypes.data, len(out), C.byref(n), None, 0, C.byref(nl))
    assert rc == L.SIFT_OK and n.value == len(cands) and nl.value == len(low)
    ora.close()


def test_mosaic_strips_against_oracle(engine):
    """BASELINE configs[4] in small: a 2048x3072 mosaic cut into 3 row strips (per-octave seed-halo exchange, walk
    hand-over), every strip's levels and the merged records against the ORACLE of the whole image -- not against
    the whole-image GPU run."""
    from sift_b200 import mosaic
    from parity import DOG_FLOOR, LEVEL_RTOL
    w, h, n_oct = 2048, 3072, 4
    u8 = fixtures.synthetic_u8(w, h, 4321)
    prm = L.default_params(numberOfOctaves=n_oct, minBlurLevel=1.6)
    ora = oracle.detect(fixtures.to_float(u8), numberOfOctaves=n_oct, minBlurLevel=1.6, separable=True)
    engines = [sift_b200.Engine(0) for _ in range(3)]
    try:
        kps, stats, lays = mosaic.detect_mosaic_local(engines, u8, prm, margin=16)
        check_keypoints(kps, ora)
        assert sum(s["keypoints"] for s in stats) == len(kps) > 10000
        worst = 0.0
        for eng, lay in zip(engines, lays):                       # owned rows of every DoG level, all octaves
            for o in range(n_oct):
                a, b = lay.own0[o] - lay.top[o], lay.own1[o] - lay.top[o]
                for s in range(5):
                    d = eng.get_level(L.SIFT_LEVEL_DOG, o, s)[a:b].astype(np.float64)
                    ref = ora.dog[o][s][lay.own0[o]:lay.own1[o]]
                    worst = max(worst, float((np.abs(d - ref) / np.maximum(np.abs(ref), DOG_FLOOR)).max()))
        assert worst <= LEVEL_RTOL, f"strip DoG rel err {worst:.3g}"
    finally:
        for e in engines:
            e.close()
        ora.close()
