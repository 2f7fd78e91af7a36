"""The host mirror of the reference's interface (sift_b200.sift == src/sift.js, sift_b200.background ==
background.js stage functions) against outputs of the reference itself (tests/golden, see make_golden.py).
These call the CUDA engine through the C ABI; they read like tests of the reference's own exports would."""
import os

import numpy as np
import pytest

import sift_b200
from sift_b200 import (SIFT_blurMatrix2DChunk, SIFT_findExtremas, SIFT_generateGradientVector,
                       SIFT_generateHessianMatrix, SIFT_subtractMatrix2DChunk, computeDifferenceOfGaussians,
                       computeGaussianScaleSpace, findCandidateKeypoints, refineCandidateKeypoints)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def steps():
    return np.load(os.path.join(GOLDEN, "ref_steps.npz"), allow_pickle=True)


def test_blur_chunk_mutates_output_and_returns_chunk(engine, steps):
    """sift.js:72-149: writes only the half-open chunk of `output`, returns the chunk matrix."""
    img = steps["blur_in"]
    output = [[0.0] * img.shape[1] for _ in range(img.shape[0])]          # Matrix2D = list of rows
    chunk = SIFT_blurMatrix2DChunk(img.tolist(), output, float(steps["blur_sigma"]),
                                   {"x1": 2, "y1": 1, "x2": 8, "y2": 10}, engine=engine)
    assert isinstance(chunk, list) and len(chunk) == 9 and len(chunk[0]) == 6
    assert np.abs(np.array(chunk) - steps["blur_chunk"]).max() <= 1e-14
    out = np.array(output)
    assert np.abs(out - steps["blur_output"]).max() <= 1e-14
    assert not out[0].any() and not out[:, :2].any() and not out[10:].any() and not out[:, 8:].any()


def test_blur_chunk_empty_chunk_is_a_no_op(engine, steps):
    img = steps["blur_in"]
    out = np.full_like(img, 7.0)
    chunk = SIFT_blurMatrix2DChunk(img, out, 1.3, {"x1": 3, "y1": 3, "x2": 3, "y2": 5}, engine=engine)
    assert chunk.size == 0 and (out == 7.0).all()


def test_subtract_chunk(engine, steps):
    """sift.js:154-188: input_pair[0] - input_pair[1] (finer minus coarser), exact in float64."""
    out = np.zeros_like(steps["sub_a"])
    chunk = SIFT_subtractMatrix2DChunk([steps["sub_a"], steps["sub_b"]], out, {"x1": 1, "y1": 0, "x2": 7, "y2": 5},
                                       engine=engine)
    assert np.array_equal(chunk, steps["sub_chunk"]) and np.array_equal(out, steps["sub_output"])


def test_find_extremas(engine, steps):
    """sift.js:212-316: strict 26-neighbour extrema in raster order, split by 0.8 * threshold."""
    res = SIFT_findExtremas(list(steps["ext_trio"]), 3, engine=engine)
    cand = np.array([(e["x"], e["y"], e["value"]) for e in res["candidateKeypoints"]]).reshape(-1, 3)
    low = np.array([(e["x"], e["y"], e["value"]) for e in res["lowContrastKeypoints"]]).reshape(-1, 3)
    assert np.array_equal(cand, steps["ext_cand"]) and np.array_equal(low, steps["ext_low"])


def test_gradient_and_hessian(engine, steps):
    """sift.js:333-353, 377-447; `dog` is indexed [o][s].image like the reference's."""
    dog = [[{"image": t} for t in steps["ext_trio"]]]
    g = SIFT_generateGradientVector(0, 1, 4, 5, dog, engine=engine)
    h = SIFT_generateHessianMatrix(0, 1, 4, 5, dog, engine=engine)
    assert np.array_equal(np.asarray(g), steps["grad"]) and np.array_equal(np.asarray(h), steps["hess"])


def test_linear_resize(engine, steps):
    """matrix2d.js:112-138."""
    assert np.array_equal(engine.linear_resize(steps["resize_in"], 0.5), steps["resize_half"])
    assert np.array_equal(engine.linear_resize(steps["resize_in"], 2.0), steps["resize_two"])
    with pytest.raises(sift_b200.SiftError):
        engine.linear_resize(steps["resize_in"], 0.75)       # non-dyadic rates are refused, not approximated


@pytest.mark.parametrize("name", ["g22x18_o2_b16", "g28x22_o2_b08"])
def test_stage_functions_reproduce_the_reference_replies(engine, name):
    """The four messages of main.js:111 -> 239 -> 274 -> 325 through the mirror's stage functions."""
    g = np.load(os.path.join(GOLDEN, f"ref_{name}.npz"), allow_pickle=True)
    n_oct, spo, min_blur, assumed = g["params"]
    n_oct, spo = int(n_oct), int(spo)
    ss = computeGaussianScaleSpace(g["input_matrix"], number_of_octaves=n_oct, scales_per_octave=spo,
                                   min_blur_level=float(min_blur), assumed_blur=float(assumed), chunk_size=32,
                                   engine=engine)
    assert len(ss) == n_oct and len(ss[0]) == spo + 3
    for o in range(n_oct):
        for s in range(spo + 3):
            ref = g[f"gauss_{o}_{s}"]
            assert abs(ss[o][s]["blurLevel"] - float(g[f"gauss_blur_{o}_{s}"])) <= 1e-12
            assert np.abs(ss[o][s]["image"] - ref).max() <= 1e-5 * np.abs(ref).max()
    dog = computeDifferenceOfGaussians(ss, engine=engine)
    for o in range(n_oct):
        assert len(dog[o]) == spo + 2
        for s in range(spo + 2):
            ref = g[f"dog_{o}_{s}"]
            assert abs(dog[o][s]["blurLevel"] - float(g[f"dog_blur_{o}_{s}"])) <= 1e-12
            assert (np.abs(dog[o][s]["image"] - ref) / np.maximum(np.abs(ref), 0.012)).max() <= 1e-5
    cands = findCandidateKeypoints(dog, [octave[0]["image"] for octave in ss], spo, engine=engine)
    got = [(o, grp["scaleLevel"], e["x"], e["y"]) for o, octave in enumerate(cands) for grp in octave
           for e in grp["localExtremas"]]
    assert got == [tuple(int(v) for v in row[:4]) for row in g["candidates"]]
    kps = refineCandidateKeypoints(dog, spo, n_oct, cands, float(min_blur), engine=engine)
    ref_kp = g["keypoints"]
    assert len(kps) == len(ref_kp)
    for k, r in zip(kps, ref_kp):
        assert (k["octave"], k["scaleLevel"], k["localX"], k["localY"]) == tuple(int(v) for v in r[:4])
        assert abs(k["absoluteX"] - r[5]) <= 1e-3 and abs(k["absoluteY"] - r[6]) <= 1e-3
        assert abs(k["absoluteSigma"] - r[4]) <= 1e-3 * r[4]


def test_foreign_dog_levels_can_be_refined(engine):
    """refineCandidateKeypoints on DoG matrices the engine did not build (plain nested lists, like a structured
    clone arriving from another worker): they are uploaded and refined against."""
    g = np.load(os.path.join(GOLDEN, "ref_g16x16_o2_s2_b10.npz"), allow_pickle=True)
    n_oct, spo = int(g["params"][0]), int(g["params"][1])
    dog = [[{"blurLevel": float(g[f"dog_blur_{o}_{s}"]), "image": g[f"dog_{o}_{s}"].tolist()} for s in range(spo + 2)]
           for o in range(n_oct)]
    cands = [[{"scaleLevel": s, "localExtremas": []} for s in range(1, spo + 1)] for _ in range(n_oct)]
    for o, s, x, y, v in g["candidates"]:
        cands[int(o)][int(s) - 1]["localExtremas"].append({"x": int(x), "y": int(y), "value": float(v)})
    kps = refineCandidateKeypoints(dog, spo, n_oct, cands, float(g["params"][2]), engine=engine)
    ref_kp = g["keypoints"]
    assert len(kps) == len(ref_kp)
    for k, r in zip(kps, ref_kp):
        assert abs(k["absoluteX"] - r[5]) <= 1e-3 and abs(k["absoluteY"] - r[6]) <= 1e-3
