"""The oracle (and, on the GPU box, the CUDA engine) against outputs of the REFERENCE ITSELF.

tests/golden/ref_*.npz were produced by oracle/make_golden.py, which runs the unmodified reference
JavaScript (background.js + src/*.js read from /root/reference) through oracle/jsmini.py and drives the
worker the way main.js does.  This is what pins the float64 C restatement in oracle/sift_oracle.c:
every Gaussian level, DoG level, candidate and refined keypoint of the reference, bit for bit.
"""
import glob
import os

import numpy as np
import pytest

import oracle
from sift_b200 import _lib as L

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.basename(p)[4:-4] for p in glob.glob(os.path.join(GOLDEN, "ref_g*.npz")))


def _load(name):
    return np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=True)


def test_golden_fixtures_exist():
    assert len(CASES) >= 4, "run oracle/make_golden.py in the build container"
    assert os.path.exists(os.path.join(GOLDEN, "ref_steps.npz"))


@pytest.mark.parametrize("name", CASES)
def test_oracle_reproduces_the_reference_bit_for_bit(name):
    g = _load(name)
    n_oct, spo, min_blur, assumed = g["params"]
    n_oct, spo = int(n_oct), int(spo)
    r = oracle.detect(g["input_matrix"], numberOfOctaves=n_oct, scalesPerOctave=spo, minBlurLevel=float(min_blur),
                      assumedBlur=float(assumed), separable=False)
    assert int(g["n_octaves"]) == n_oct and int(g["n_levels"]) == spo + 3
    for o in range(n_oct):
        for s in range(spo + 3):
            ref = g[f"gauss_{o}_{s}"]
            assert r.gauss[o][s].shape == ref.shape
            assert np.array_equal(r.gauss[o][s], ref), f"gauss {o},{s}: max diff {np.abs(r.gauss[o][s] - ref).max():.3g}"
            assert r.blur[o][s] == float(g[f"gauss_blur_{o}_{s}"])
        for s in range(spo + 2):
            assert np.array_equal(r.dog[o][s], g[f"dog_{o}_{s}"])
    cand = np.array([(c["octave"], c["scale"], c["x"], c["y"], c["value"]) for c in r.candidates]).reshape(-1, 5)
    assert np.array_equal(cand, g["candidates"])
    assert r.n_low_contrast == int(g["n_low_contrast_markers"])
    kp = np.array([(k["octave"], k["scaleLevel"], k["localX"], k["localY"], k["absoluteSigma"], k["absoluteX"],
                    k["absoluteY"], k["interpolatedValue"]) for k in r.keypoints]).reshape(-1, 8)
    ref_kp = g["keypoints"]
    assert kp.shape == ref_kp.shape and len(ref_kp) > 0
    assert np.array_equal(kp[:, :4], ref_kp[:, :4])
    # positions / value: identical operation order -> identical doubles; sigma goes through pow (libm vs V8: 1 ulp)
    assert np.array_equal(kp[:, 5:8], ref_kp[:, 5:8])
    assert np.allclose(kp[:, 4], ref_kp[:, 4], rtol=4e-16, atol=0)


def test_reference_results_do_not_depend_on_chunk_size():
    """background.js:147-203 blurs in chunk_size tiles only to repaint progressively; every tile reads the whole base
    image (sift.js:109-125).  The unmodified reference, run with chunk sizes 32 and 5 (11 vs 225 chunk messages),
    produced identical levels, DoG, candidates and keypoints -- which is why the C ABI has no chunk parameter."""
    g = np.load(os.path.join(GOLDEN, "ref_chunking.npz"))
    assert int(g["c5_chunk_messages"]) > 10 * int(g["c32_chunk_messages"]) > 0
    for key in ("levels", "dog1", "candidates", "keypoints"):
        assert np.array_equal(g[f"c32_{key}"], g[f"c5_{key}"]), key
    r = oracle.detect(g["input_u8"].astype(np.float64) / 255.0, numberOfOctaves=2, scalesPerOctave=3, minBlurLevel=1.0,
                      assumedBlur=0.5, separable=False)
    assert np.array_equal(np.stack(r.gauss[0]), g["c5_levels"]) and np.array_equal(np.stack(r.dog[1]), g["c5_dog1"])


def test_rgba_ingest_matches_image_utils():
    """image-utils.js:107-114: grey = (0.299R + 0.587G + 0.114B) / 255, which is NOT v/255 for grey bytes."""
    g = _load("g17x13_o3_b08_rgba")
    u8 = g["input_u8"].astype(np.float64)
    want = ((u8 * 0.299) + (u8 * 0.587) + (u8 * 0.114)) / 255.0
    assert np.array_equal(g["input_matrix"], want)
    assert not np.array_equal(g["input_matrix"], u8 / 255.0)


def test_oracle_step_functions_match_the_reference_exports():
    """Direct calls of src/sift.js's five exports and the matrix2d helpers (make_golden.step_function_vectors)."""
    g = np.load(os.path.join(GOLDEN, "ref_steps.npz"), allow_pickle=True)
    img = g["blur_in"]
    out = np.zeros_like(img)
    chunk = oracle.blur_chunk(img, out, float(g["blur_sigma"]), 2, 1, 8, 10)
    assert np.array_equal(chunk, g["blur_chunk"]) and np.array_equal(out, g["blur_output"])
    out = np.zeros_like(g["sub_a"])
    chunk = oracle.subtract_chunk(g["sub_a"], g["sub_b"], out, 1, 0, 7, 5)
    assert np.array_equal(chunk, g["sub_chunk"]) and np.array_equal(out, g["sub_output"])
    res = oracle.find_extremas(list(g["ext_trio"]), 3)
    cand = np.array([(e["x"], e["y"], e["value"]) for e in res["candidateKeypoints"]]).reshape(-1, 3)
    low = np.array([(e["x"], e["y"], e["value"]) for e in res["lowContrastKeypoints"]]).reshape(-1, 3)
    assert len(g["ext_cand"]) >= 2
    assert np.array_equal(cand, g["ext_cand"]) and np.array_equal(low, g["ext_low"])
    assert np.array_equal(oracle.gradient(list(g["ext_trio"]), 1, 4, 5), g["grad"])
    assert np.array_equal(oracle.hessian(list(g["ext_trio"]), 1, 4, 5), g["hess"])
    assert np.array_equal(oracle.inverse3x3(g["inv_in"]), g["inv_out"])
    assert bool(g["inv_singular_is_null"]) and oracle.inverse3x3(np.array([[1., 2, 3], [2, 4, 6], [1, 1, 1]])) is None
    assert np.array_equal(oracle.linear_resize(g["resize_in"], 0.5), g["resize_half"])
    assert np.array_equal(oracle.linear_resize(g["resize_in"], 2.0), g["resize_two"])
    # image-utils.js:295-332: x-outer / y-inner chunk order
    b = g["chunk_bounds_70x45"]
    assert [tuple(r) for r in b] == [(0, 0, 32, 32), (0, 32, 32, 45), (32, 0, 64, 32), (32, 32, 64, 45),
                                     (64, 0, 70, 32), (64, 32, 70, 45)]


# ------------------------------------------------------------------ the CUDA engine against the reference
@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_matches_the_reference(engine, name):
    """North-star tolerances: levels within 1e-5 relative, keypoints on the same cell within 1e-3 px."""
    g = _load(name)
    n_oct, spo, min_blur, assumed = g["params"]
    n_oct, spo = int(n_oct), int(spo)
    prm = L.default_params(numberOfOctaves=n_oct, scalesPerOctave=spo, minBlurLevel=float(min_blur),
                           assumedBlur=float(assumed))
    if str(g["ingest"]) == "rgba":
        u8 = g["input_u8"]
        rgba = np.repeat(u8[:, :, None], 4, axis=2).copy()
        rgba[:, :, 3] = 255
        engine.build_scale_space(rgba, prm, rgba=True)
        kps, _ = engine.detect(rgba, prm, rgba=True)
    else:
        engine.build_scale_space(g["input_u8"], prm)
        kps, _ = engine.detect(g["input_u8"], prm)
    for o in range(n_oct):
        for s in range(spo + 3):
            got = engine.get_level(L.SIFT_LEVEL_GAUSSIAN, o, s).astype(np.float64)
            ref = g[f"gauss_{o}_{s}"]
            assert np.abs(got - ref).max() <= 1e-5 * np.abs(ref).max()
        for s in range(spo + 2):
            got = engine.get_level(L.SIFT_LEVEL_DOG, o, s).astype(np.float64)
            ref = g[f"dog_{o}_{s}"]
            assert (np.abs(got - ref) / np.maximum(np.abs(ref), 0.012)).max() <= 1e-5
    cands, _ = engine.find_candidates()
    got_c = [(int(c["octave"]), int(c["scaleLevel"]), int(c["x"]), int(c["y"])) for c in cands]
    assert got_c == [tuple(int(v) for v in row[:4]) for row in g["candidates"]]
    ref_kp = g["keypoints"]
    assert len(kps) == len(ref_kp)
    for k, r in zip(kps, ref_kp):
        assert (int(k["octave"]), int(k["scaleLevel"]), int(k["localX"]), int(k["localY"])) == tuple(int(v) for v in r[:4])
        assert abs(float(k["absoluteX"]) - r[5]) <= 1e-3 and abs(float(k["absoluteY"]) - r[6]) <= 1e-3
        assert abs(float(k["absoluteSigma"]) - r[4]) <= 1e-3 * r[4]
        assert abs(float(k["interpolatedValue"]) - r[7]) <= 1e-5 * max(abs(r[7]), 0.012)
