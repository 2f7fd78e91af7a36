"""Known-answer tests of the float64 oracle (CPU only).

The reference ships no tests, golden vectors or fixtures, so the oracle is pinned
against analytic answers derived from the cited reference lines (KAT 1-10) and against
independent numpy restatements written here from the same lines.
"""
import math

import numpy as np
import pytest

import oracle
from sift_b200 import fixtures


def test_kat1_constant_image():
    """Constant image: every level equals the constant (normalised kernel, sift.js:48-63), no extrema."""
    img = np.full((24, 31), 0.37)
    r = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6)
    for o in range(3):
        for s in range(6):
            assert np.abs(r.gauss[o][s] - 0.37).max() < 1e-14
    assert r.candidates == [] and r.keypoints == [] and r.n_low_contrast == 0


def test_kat2_impulse_is_outer_product_of_1d_kernel():
    """sift.js:22-67: g(i,j)/sum is the outer product of the normalised 1D Gaussian, R = round(3 sigma)."""
    sigma = 1.9529
    R = oracle.kernel_radius(sigma)
    assert R == 6
    k = oracle.build_gaussian_kernel(sigma)
    assert k.shape == (13, 13)
    assert abs(k.sum() - 1.0) < 1e-14
    x = np.arange(-R, R + 1, dtype=np.float64)
    w = np.exp(-0.5 * x * x / sigma ** 2)
    w /= w.sum()
    assert np.abs(k - np.outer(w, w)).max() < 1e-16
    img = np.zeros((41, 41))
    img[20, 20] = 1.0
    out = oracle.blur_image(img, sigma)
    assert np.abs(out[20 - R:20 + R + 1, 20 - R:20 + R + 1] - k).max() < 1e-18
    out[20 - R:20 + R + 1, 20 - R:20 + R + 1] = 0
    assert not out.any()                                # support is exactly (2R+1)^2


def test_kat3_narrow_image_equals_edge_padded_correlation():
    """sift.js:109-119: sample coordinates clamp per axis == numpy pad(mode='edge')."""
    rng = np.random.default_rng(3)
    img = rng.random((9, 4))
    sigma = 2.4901                                      # R = 7 > both dimensions
    R = oracle.kernel_radius(sigma)
    k = oracle.build_gaussian_kernel(sigma)
    pad = np.pad(img, R, mode="edge")
    want = np.empty_like(img)
    for y in range(img.shape[0]):
        for x in range(img.shape[1]):
            # kernel[i][j]: i = x offset, j = y offset (sift.js:104-125)
            want[y, x] = (pad[y:y + 2 * R + 1, x:x + 2 * R + 1] * k.T).sum()
    got = oracle.blur_image(img, sigma)
    assert np.abs(got - want).max() < 1e-15
    sep = oracle.blur_image(img, sigma, separable=True)
    assert np.abs(sep - want).max() < 1e-15


def test_kat4_linear_resize():
    """matrix2d.js:112-138: rate 0.5 duplicates 2x2, rate 2.0 takes [::2, ::2] with ceil sizes."""
    m = np.arange(35, dtype=np.float64).reshape(5, 7)
    up = oracle.linear_resize(m, 0.5)
    assert up.shape == (10, 14)
    assert np.array_equal(up, np.repeat(np.repeat(m, 2, axis=0), 2, axis=1))
    dn = oracle.linear_resize(m, 2.0)
    assert dn.shape == (3, 4)
    assert np.array_equal(dn, m[::2, ::2])


def test_kat5_dog_sign_finer_minus_coarser():
    """sift.js:172: D = S[s-1] - S[s]; a bright blob gives a POSITIVE DoG maximum at its centre."""
    yy, xx = np.mgrid[0:48, 0:48]
    img = 0.2 + 0.6 * np.exp(-((xx - 24.3) ** 2 + (yy - 23.6) ** 2) / (2 * 1.6 ** 2))
    r = oracle.detect(img, numberOfOctaves=1, minBlurLevel=1.6)
    for s in range(5):
        assert np.abs(r.dog[0][s] - (r.gauss[0][s] - r.gauss[0][s + 1])).max() == 0.0
    assert r.dog[0][1][48, 49] > 0
    assert [(c["scale"], c["x"], c["y"]) for c in r.candidates] == [(3, 49, 48)] and r.candidates[0]["value"] > 0


def test_kat5b_pixel_centred_blob_is_a_rounding_noise_tie():
    """Q3 + sift.js:261,266: the nearest-neighbour upsample turns a pixel-centred symmetric blob into a 2x2
    plateau whose four DoG values agree to rounding noise; strict comparisons let at most one of them through,
    and WHICH one is decided by the last bits -- the 'exact-tie neighbourhood' class of allowed mismatches."""
    yy, xx = np.mgrid[0:48, 0:48]
    img = 0.2 + 0.6 * np.exp(-((xx - 24) ** 2 + (yy - 24) ** 2) / (2 * 1.6 ** 2))
    r = oracle.detect(img, numberOfOctaves=1, minBlurLevel=1.6)
    quad = r.dog[0][3][48:50, 48:50]
    assert quad.min() > 0.06 and quad.max() - quad.min() < 1e-14
    assert len([c for c in r.candidates if c["scale"] == 3 and c["x"] in (48, 49) and c["y"] in (48, 49)]) <= 1


def test_kat6_plateau_is_not_an_extremum():
    """sift.js:261,266: strict comparisons, ties are never extrema."""
    d0 = np.zeros((5, 6)); d2 = np.zeros((5, 6)); d1 = np.zeros((5, 6))
    d1[2, 2] = 1.0
    res = oracle.find_extremas([d0, d1, d2])
    assert [(e["x"], e["y"]) for e in res["candidateKeypoints"]] == [(2, 2)]
    d1[2, 3] = 1.0                                      # two equal neighbours: neither is strict
    res = oracle.find_extremas([d0, d1, d2])
    assert res["candidateKeypoints"] == [] and res["lowContrastKeypoints"] == []
    d1[2, 3] = 0.0
    d1[2, 2] = 0.011                                    # below 0.8 * 0.015 -> low-contrast list (sift.js:293-306)
    res = oracle.find_extremas([d0, d1, d2])
    assert res["candidateKeypoints"] == [] and len(res["lowContrastKeypoints"]) == 1
    d1[2, 2] = -1.0                                     # strict minimum
    res = oracle.find_extremas([d0, d1, d2])
    assert [(e["x"], e["y"], e["value"]) for e in res["candidateKeypoints"]] == [(2, 2, -1.0)]


def test_kat6b_border_pixels_are_never_scanned():
    """sift.js:221-222: y in [1, h-2], x in [1, w-2]."""
    d0 = np.zeros((4, 4)); d2 = np.zeros((4, 4)); d1 = np.zeros((4, 4))
    d1[0, 0] = d1[3, 3] = d1[0, 2] = 5.0
    assert oracle.find_extremas([d0, d1, d2])["candidateKeypoints"] == []


def test_kat7_quadratic_step_and_singular_inverse():
    """matrix2d.js:464-509: cofactor inverse; abs(det) < Number.EPSILON -> null."""
    H = np.diag([-2.0, -2.0, -2.0])
    inv = oracle.inverse3x3(H)
    g = np.array([0.2, -0.4, 0.6])
    alpha = (-inv) @ g
    assert np.allclose(alpha, [0.1, -0.2, 0.3], rtol=0, atol=1e-16)
    assert oracle.inverse3x3(np.zeros((3, 3))) is None
    assert oracle.inverse3x3(np.diag([1e-6, 1e-6, 1e-5])) is None        # det 1e-17 < 2.22e-16
    assert oracle.inverse3x3(np.diag([1e-5, 1e-5, 1e-5])) is not None    # det 1e-15
    rng = np.random.default_rng(0)
    m = rng.random((3, 3)) + np.eye(3)
    assert np.abs(oracle.inverse3x3(m) - np.linalg.inv(m)).max() < 1e-13


def test_kat7b_gradient_and_hessian_stencils():
    """sift.js:333-353 and 377-447 on a quadratic form: derivatives are exact for a quadratic."""
    A = np.array([[1.0, 0.25, -0.5], [0.25, -2.0, 0.75], [-0.5, 0.75, 3.0]])   # [s, m, n]
    b = np.array([0.3, -0.2, 0.1])
    f = lambda s, m, n: 0.5 * np.array([s, m, n]) @ A @ np.array([s, m, n]) + b @ np.array([s, m, n])
    dog = [np.array([[f(s - 1, m - 2, n - 2) for n in range(5)] for m in range(5)]) for s in range(3)]
    g = oracle.gradient(dog, 1, 2, 2)
    h = oracle.hessian(dog, 1, 2, 2)
    assert np.allclose(g, b, atol=1e-14)
    assert np.allclose(h, A, atol=1e-13)


def test_kat8_threshold_constants():
    """sift.js:285-293, background.js:572, 598."""
    thr = oracle.lib().oracle_contrast_threshold(3, 0.015)
    assert abs(thr - 0.015) < 1e-17
    assert abs(0.8 * thr - 0.012) < 1e-17
    assert abs((10 + 1) ** 2 / 10 - 12.1) < 1e-14
    thr5 = oracle.lib().oracle_contrast_threshold(5, 0.015)
    assert abs(thr5 - 0.015 * (2 ** (1 / 5) - 1) / (2 ** (1 / 3) - 1)) < 1e-17


def test_kat9_absolute_coordinates():
    """background.js:611-614: delta = 2^(octave-1); absolute = delta * (alpha + index)."""
    yy, xx = np.mgrid[0:64, 0:64]
    img = 0.2 + 0.6 * np.exp(-((xx - 30.3) ** 2 + (yy - 21.4) ** 2) / (2 * 1.6 ** 2))
    r = oracle.detect(img, numberOfOctaves=2, minBlurLevel=1.6)
    assert r.keypoints
    for k in r.keypoints:
        delta = 2.0 ** (k["octave"] - 1)
        assert k["absoluteX"] == delta * (k["offset"][2] + k["localX"])
        assert k["absoluteY"] == delta * (k["offset"][1] + k["localY"])
        want = (delta / 0.5) * 1.6 * 2.0 ** ((k["offset"][0] + k["scaleLevel"]) / 3)
        assert abs(k["absoluteSigma"] - want) <= 1e-15 * want
        assert all(abs(a) < 0.6 for a in k["offset"])
    best = max(r.keypoints, key=lambda k: abs(k["interpolatedValue"]))
    # octave-0 pixel X covers source [X/2, X/2 + 0.5): its centre is X/2 + 0.25 in source pixels
    assert abs(best["absoluteX"] - 30.55) < 0.02 and abs(best["absoluteY"] - 21.65) < 0.02


SIGMA_TABLE = {   # offset sigma / radius per octave for minBlurLevel 1.6, assumedBlur 0.5, spo 3
    0: [(1.5199, 5), (1.9529, 6), (2.4901, 7), (3.1607, 9), (4.0006, 12), (5.0550, 15)],
    1: [None, (2.4525, 7), (3.9450, 12), (5.5426, 17), (7.4013, 22), (9.6422, 29)],
    2: [None, (4.9051, 15), (7.8900, 24), (11.0851, 33), (14.8027, 44), (19.2845, 58)],
    3: [None, (9.8102, 29), (15.7801, 47), (22.1703, 67), (29.6054, 89), (38.5689, 116)],
}


def test_kat10_sigma_schedule_and_radii():
    """background.js:156-177 + sift.js:38 (Q1: the blur level doubles per octave in LOCAL pixels)."""
    img = fixtures.to_float(fixtures.synthetic_u8(40, 40, 1))
    r = oracle.detect(img, numberOfOctaves=4, minBlurLevel=1.6, separable=True)
    for o, row in SIGMA_TABLE.items():
        assert abs(r.blur[o][0] - 1.6 * 2 ** o) < 1e-12
        for s, want in enumerate(row):
            if want is None:
                assert r.offset_sigma[o][s] == 0
                continue
            assert abs(r.offset_sigma[o][s] - want[0]) < 6e-5
            assert oracle.kernel_radius(r.offset_sigma[o][s]) == want[1]
            assert abs(r.blur[o][s] - 1.6 * 2 ** o * 2 ** (s / 3)) < 1e-12
    assert r.shapes == [(80, 80), (40, 40), (20, 20), (10, 10)]


def test_js_round_is_round_half_up():
    """Q9: Math.round(-2.5) == -2, Math.round(2.5) == 3."""
    assert oracle.js_round(-2.5) == -2 and oracle.js_round(2.5) == 3
    assert oracle.js_round(-0.5) == 0 and oracle.js_round(0.49999999999999994) == 0
    assert oracle.js_round(7.4704) == 7 and oracle.js_round(-7.51) == -8


def test_seed_is_decimated_unblurred_level_spo():
    """background.js:114-130: octave o>0 level 0 = S[o-1][spo][::2, ::2], stored unblurred."""
    img = fixtures.to_float(fixtures.synthetic_u8(37, 29, 5))
    r = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6)
    for o in (1, 2):
        assert np.array_equal(r.gauss[o][0], r.gauss[o - 1][3][::2, ::2])
        assert r.blur[o][0] == r.blur[o - 1][3]


def test_every_level_is_blurred_from_the_octave_base():
    """Q2 (background.js:185-190): level s = blur(base, offset sigma_s), never blur(level s-1)."""
    img = fixtures.to_float(fixtures.synthetic_u8(24, 20, 9))
    r = oracle.detect(img, numberOfOctaves=2, minBlurLevel=1.6)
    base0 = oracle.linear_resize(img, 0.5)                                 # background.js:84
    for s in range(6):
        assert np.array_equal(r.gauss[0][s], oracle.blur_image(base0, r.offset_sigma[0][s]))
    for s in range(1, 6):
        assert np.array_equal(r.gauss[1][s], oracle.blur_image(r.gauss[1][0], r.offset_sigma[1][s]))


def test_chunked_blur_equals_whole_image_blur():
    """background.js:181-203: the 32x32 chunk grid does not change results (each output pixel is independent)."""
    img = fixtures.to_float(fixtures.synthetic_u8(50, 41, 2))
    whole = oracle.blur_image(img, 2.0)
    out = np.zeros_like(img)
    for y1 in range(0, 41, 32):
        for x1 in range(0, 50, 32):
            oracle.blur_chunk(img, out, 2.0, x1, y1, min(x1 + 32, 50), min(y1 + 32, 41))
    assert np.array_equal(out, whole)


def test_separable_oracle_equals_dense_oracle():
    """The separable float64 variant used for the large parity cases equals the dense 2D kernel to ~1e-15."""
    img = fixtures.to_float(fixtures.synthetic_u8(48, 40, 17))
    a = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6, separable=False)
    b = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6, separable=True)
    for o in range(3):
        for s in range(6):
            assert np.abs(a.gauss[o][s] - b.gauss[o][s]).max() < 2e-14
    key = lambda k: (k["octave"], k["candScale"], k["candY"], k["candX"])
    assert [key(k) for k in a.keypoints] == [key(k) for k in b.keypoints]


def test_separable_oracle_equals_dense_oracle_128x96():
    """before the separable variant stands in for the dense kernel at 1080p / 4K, it must equal it
    to ~1e-14 on a case both can run -- levels, DoG, the candidate and low-contrast lists, and the keypoints."""
    img = fixtures.to_float(fixtures.synthetic_u8(128, 96, 1234))
    a = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6, separable=False)
    b = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6, separable=True)
    for o in range(3):
        for s in range(6):
            assert np.abs(a.gauss[o][s] - b.gauss[o][s]).max() < 1e-14
        for s in range(5):
            assert np.abs(a.dog[o][s] - b.dog[o][s]).max() < 1e-14
    ck = lambda c: (c["octave"], c["scale"], c["y"], c["x"])
    assert [ck(c) for c in a.candidates] == [ck(c) for c in b.candidates] and len(a.candidates) > 50
    assert [ck(c) for c in a.low_contrast] == [ck(c) for c in b.low_contrast] and a.n_low_contrast == len(a.low_contrast) > 20
    key = lambda k: (k["octave"], k["candScale"], k["candY"], k["candX"], k["scaleLevel"], k["localX"], k["localY"])
    assert [key(k) for k in a.keypoints] == [key(k) for k in b.keypoints]
    for ka, kb in zip(a.keypoints, b.keypoints):
        assert abs(ka["absoluteX"] - kb["absoluteX"]) < 1e-9 and abs(ka["absoluteY"] - kb["absoluteY"]) < 1e-9


def test_refine_margin_diagnostic():
    """oracle_refine_margin (test diagnostic used by tests/parity.py): zero-ish only for walks that sit on a decision."""
    img = fixtures.to_float(fixtures.synthetic_u8(128, 96, 1234))
    r = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6, separable=True)
    assert set(r.margins) == {(c["octave"], c["scale"], c["y"], c["x"]) for c in r.candidates}
    assert all(m >= 0 for m in r.margins.values()) and min(r.margins.values()) > 1e-7
    # moving the contrast threshold onto a keypoint's interpolated value puts its margin at ~0
    k = r.keypoints[0]
    c = abs(k["interpolatedValue"])
    r2 = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6, separable=True, contrastThreshold=c * (1 + 1e-9),
                       preFilterFactor=0.8 * 0.015 / c)
    assert r2.margins[(k["octave"], k["candScale"], k["candY"], k["candX"])] < 1e-8


def test_refine_on_supplied_dog_levels():
    """oracle.refine_on_dog (used by tests/parity.py to tell a device's arithmetic from its float32 level storage):
    on the oracle's own DoG it reproduces the oracle's records bit for bit; on float32-rounded levels the records
    stay on the same cells with positions within the storage error."""
    img = fixtures.to_float(fixtures.synthetic_u8(128, 96, 1234))
    prm = oracle.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    r = oracle.detect(img, prm, separable=True)
    assert len(r.keypoints) > 30
    for k in r.keypoints:
        o = k["octave"]
        cand = {"scale": k["candScale"], "x": k["candX"], "y": k["candY"], "value": k["dogValue"]}
        out, rec = oracle.refine_on_dog(r.dog[o], o, cand, prm)
        assert out == "accepted" and rec == k
        f32 = [d.astype(np.float32).astype(np.float64) for d in r.dog[o]]
        if r.margins[(o, k["candScale"], k["candY"], k["candX"])] > 1e-3:
            out32, rec32 = oracle.refine_on_dog(f32, o, dict(cand, value=float(np.float32(k["dogValue"]))), prm)
            assert out32 == "accepted" and (rec32["localX"], rec32["localY"]) == (k["localX"], k["localY"])
            assert abs(rec32["absoluteX"] - k["absoluteX"]) < 1e-3


def test_refine_uses_original_value_and_keeps_duplicates():
    """Q5 (background.js:565): omega = extrema.value + 0.5 alpha.g with the ORIGINAL candidate value."""
    img = fixtures.to_float(fixtures.synthetic_u8(96, 80, 42))
    r = oracle.detect(img, numberOfOctaves=3, minBlurLevel=1.6)
    moved = [k for k in r.keypoints if k["iterations"] > 0]
    assert moved, "fixture should contain at least one keypoint that moved before converging"
    for k in r.keypoints:
        dog = r.dog[k["octave"]]
        g = oracle.gradient(dog, k["scaleLevel"], k["localY"], k["localX"])
        a = np.array(k["offset"])
        want = k["dogValue"] + (((0.5 * a[0]) * g[0]) + ((0.5 * a[1]) * g[1]) + ((0.5 * a[2]) * g[2]))
        assert k["interpolatedValue"] == want
        assert k["dogValue"] == dog[k["candScale"]][k["candY"], k["candX"]]
    assert sum(r.outcomes.values()) == len(r.candidates)
    assert r.outcomes["accepted"] == len(r.keypoints)


def test_numpy_restatement_of_the_full_path():
    """Independent restatement (numpy, written from the same reference lines) of scale space -> DoG -> scan."""
    img = fixtures.to_float(fixtures.synthetic_u8(64, 48, 8, blobs=40))
    r = oracle.detect(img, numberOfOctaves=2, minBlurLevel=1.6)

    def blur(a, sigma):
        R = int(math.floor(3 * sigma + 0.5))
        x = np.arange(-R, R + 1, dtype=np.float64)
        k2 = np.exp(-0.5 * (x[:, None] ** 2 + x[None, :] ** 2) / sigma ** 2) / (2 * math.pi * sigma ** 2)
        k2 /= k2.sum()
        pad = np.pad(a, R, mode="edge")
        out = np.zeros_like(a)
        for i in range(2 * R + 1):          # x offset
            for j in range(2 * R + 1):      # y offset
                out += pad[j:j + a.shape[0], i:i + a.shape[1]] * k2[i, j]
        return out

    k = 2 ** (1 / 3)
    base = np.repeat(np.repeat(img, 2, axis=0), 2, axis=1)
    levels0 = [blur(base, math.sqrt((1.6 * k ** s) ** 2 - 0.5 ** 2)) for s in range(6)]
    seed = levels0[3][::2, ::2]
    bl = 1.6 * k ** 3
    levels1 = [seed] + [blur(seed, math.sqrt((bl * k ** s) ** 2 - bl ** 2)) for s in range(1, 6)]
    for s in range(6):
        assert np.abs(levels0[s] - r.gauss[0][s]).max() < 1e-13
        assert np.abs(levels1[s] - r.gauss[1][s]).max() < 1e-13
    dog0 = [levels0[s] - levels0[s + 1] for s in range(5)]
    got = {(c["scale"], c["y"], c["x"]) for c in r.candidates if c["octave"] == 0}
    want = set()
    for s in (1, 2, 3):
        cube = np.stack(dog0[s - 1:s + 2])
        for y in range(1, cube.shape[1] - 1):
            for x in range(1, cube.shape[2] - 1):
                nb = np.delete(cube[:, y - 1:y + 2, x - 1:x + 2].ravel(), 13)
                c = cube[1, y, x]
                if ((nb > c).all() or (nb < c).all()) and abs(c) >= 0.8 * 0.015:
                    want.add((s, y, x))
    assert got == want and len(got) > 3
