"""Display products: the RGBA ImageData the reference posts beside its stage results.

CPU: the oracle restatement against bytes produced by the unmodified reference (tests/golden/ref_preview.npz,
written by oracle/make_golden.py --only preview through oracle/jsmini.py).  GPU: sift_get_level_preview against
the oracle on the engine's own levels, and the per-level messages of the host mirror against the message counts
the reference produced for the same case."""
import os

import numpy as np
import pytest

import oracle
import sift_b200
from sift_b200 import _lib as L, background, fixtures

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_preview_matches_the_reference_bytes():
    g = np.load(os.path.join(GOLDEN, "ref_preview.npz"))
    rgba, mm = oracle.preview(g["gray_in"], oracle.PREVIEW_GRAY)
    assert np.array_equal(rgba, g["gray_rgba"]) and mm == (0.0, 1.0)
    assert list(rgba[0, :6, 0]) == [0, 255, 128, 128, 255, 0]       # Math.round ties up; Uint8ClampedArray clamps
    assert (rgba[..., 3] == 255).all() and np.array_equal(rgba[..., 0], rgba[..., 1]) and np.array_equal(rgba[..., 0], rgba[..., 2])
    rgba, _ = oracle.preview(g["dog_in"], oracle.PREVIEW_SIGMOID, 5.0)                # background.js:303-307
    assert np.array_equal(rgba, g["sigmoid_rgba"])
    rgba, mm = oracle.preview(g["dog_in"], oracle.PREVIEW_MINMAX)                     # background.js:336
    assert np.array_equal(rgba, g["minmax_rgba"])
    assert mm == (float(g["dog_in"].min()), float(g["dog_in"].max()))
    assert rgba[..., 0].min() == 0 and rgba[..., 0].max() == 255


def test_oracle_preview_degenerate_inputs():
    flat = np.full((3, 4), 0.25)
    rgba, _ = oracle.preview(flat, oracle.PREVIEW_MINMAX)          # (v - min) / 0 = NaN -> Uint8ClampedArray stores 0
    assert (rgba[..., 0] == 0).all() and (rgba[..., 3] == 255).all()
    rgba, _ = oracle.preview(np.array([[0.0]]), oracle.PREVIEW_SIGMOID, 5.0)
    assert rgba[0, 0, 0] == 128                                     # 0.5 * 255 = 127.5 -> Math.round -> 128


@pytest.mark.gpu
def test_engine_previews_match_the_oracle(engine):
    u8 = fixtures.synthetic_u8(150, 110, 5)
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    engine.build_scale_space(u8, prm)
    engine.build_dog()
    for kind, o, s, mode, coef in ((L.SIFT_LEVEL_GAUSSIAN, 0, 2, L.SIFT_PREVIEW_GRAY, 1.0),
                                   (L.SIFT_LEVEL_GAUSSIAN, 2, 0, L.SIFT_PREVIEW_GRAY, 1.0),
                                   (L.SIFT_LEVEL_DOG, 0, 1, L.SIFT_PREVIEW_MINMAX, 1.0),
                                   (L.SIFT_LEVEL_DOG, 1, 4, L.SIFT_PREVIEW_MINMAX, 1.0),
                                   (L.SIFT_LEVEL_DOG, 1, 3, L.SIFT_PREVIEW_SIGMOID, 5.0)):
        lvl = engine.get_level(kind, o, s)
        got, mm = engine.level_preview(kind, o, s, mode, coef)
        want, wmm = oracle.preview(lvl.astype(np.float64), mode, coef)
        assert got.shape == want.shape == lvl.shape + (4,)
        if mode == L.SIFT_PREVIEW_SIGMOID:
            # exp() of the device vs libm may differ in the last place: at most one grey level, almost nowhere
            d = np.abs(got.astype(int) - want.astype(int))
            assert d.max() <= 1 and (d > 0).mean() < 1e-3
        else:
            assert np.array_equal(got, want)
        assert mm == wmm
    with pytest.raises(sift_b200.SiftError):
        engine.level_preview(L.SIFT_LEVEL_DOG, 0, 1, 7)


@pytest.mark.gpu
def test_stage_mirror_posts_the_reference_level_images(engine):
    """Same case as tests/golden/ref_g22x18_o2_b16.npz: the reference posted 12 Gaussian and 10 DoG level images."""
    g = np.load(os.path.join(GOLDEN, "ref_g22x18_o2_b16.npz"), allow_pickle=True)
    counts = {k: int(v) for k, v in g["msgcount_scale_space"].tolist() + g["msgcount_dog"].tolist()}
    n_oct, spo, min_blur, assumed = g["params"]
    msgs = []
    ss = background.computeGaussianScaleSpace(g["input_matrix"], int(n_oct), int(spo), float(min_blur), float(assumed),
                                              engine=engine, post_message=msgs.append)
    background.computeDifferenceOfGaussians(ss, engine=engine, post_message=msgs.append)
    by_type = {}
    for m in msgs:
        by_type.setdefault(m["type"], []).append(m)
    assert len(by_type[background.RECEIVED_GAUSSIAN_BLURRED_IMAGE]) == counts["received-gaussian-blurred-image"]
    assert len(by_type[background.RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE]) == counts["received-difference-of-gaussian-image"]
    # payloads: the reference's own fp64 levels through the oracle's ImageData restatement (fp32 storage here:
    # a byte may differ by one where v * 255 sits on a rounding boundary)
    k = 0
    for o in range(int(n_oct)):
        for s in range(int(spo) + 3):
            m = by_type[background.RECEIVED_GAUSSIAN_BLURRED_IMAGE][k]; k += 1
            want, _ = oracle.preview(g[f"gauss_{o}_{s}"], oracle.PREVIEW_GRAY)
            assert m["octave"] == o and m["imageData"]["data"].shape == want.shape
            assert np.abs(m["imageData"]["data"].astype(int) - want.astype(int)).max() <= 1
    k = 0
    for o in range(int(n_oct)):
        for s in range(int(spo) + 2):
            m = by_type[background.RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE][k]; k += 1
            want, _ = oracle.preview(g[f"dog_{o}_{s}"], oracle.PREVIEW_MINMAX)
            assert m["octave"] == o
            assert np.abs(m["imageData"]["data"].astype(int) - want.astype(int)).max() <= 1
    chunk = background.dogChunkPreview(0, 1, {"x1": 4, "y1": 2, "x2": 20, "y2": 12}, engine=engine)
    want, _ = oracle.preview(g["dog_0_1"], oracle.PREVIEW_SIGMOID, 5.0)
    assert chunk["dx"] == 4 and chunk["dy"] == 2 and chunk["imageData"]["data"].shape == (10, 16, 4)
    assert np.abs(chunk["imageData"]["data"].astype(int) - want[2:12, 4:20].astype(int)).max() <= 1
