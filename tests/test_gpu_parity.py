"""GPU parity: the CUDA path (through the C ABI / host mirror) against the float64 oracle."""
import numpy as np
import pytest

import oracle
import sift_b200
from sift_b200 import _lib as L, fixtures
from parity import check_candidates, check_keypoints, check_levels

pytestmark = pytest.mark.gpu

CASES = [
    # (width, height, octaves, seed, dense oracle?)
    (128, 96, 3, 1234, True),
    (97, 61, 3, 7, True),          # odd sizes: ceil halving (matrix2d.js:119)
    (64, 64, 4, 99, True),
    (256, 192, 4, 5, False),
    (512, 512, 4, 1234, False),    # BASELINE configs[0]
]


def _run_case(engine, w, h, n_oct, seed, dense, min_blur=1.6, dtype="u8"):
    u8 = fixtures.synthetic_u8(w, h, seed)
    img = fixtures.to_float(u8)
    ora = oracle.detect(img, numberOfOctaves=n_oct, minBlurLevel=min_blur, separable=not dense)
    assert ora.outcomes["singular"] == 0, "fixture would crash the reference (SURVEY.md Q7)"
    prm = L.default_params(numberOfOctaves=n_oct, minBlurLevel=min_blur)
    src = {"u8": u8, "f32": img.astype(np.float32), "f64": img}[dtype]
    engine.build_scale_space(src, prm)
    if dtype != "f32":
        check_levels(engine, ora, L)
    cands, _ = engine.find_candidates()
    kps, stats = engine.detect(src, prm)
    return ora, cands, kps, stats


@pytest.mark.parametrize("w,h,n_oct,seed,dense", CASES)
def test_detect_matches_oracle(engine, w, h, n_oct, seed, dense):
    ora, cands, kps, stats = _run_case(engine, w, h, n_oct, seed, dense)
    check_candidates(cands, ora)
    matched, total, worst = check_keypoints(kps, ora)
    assert stats["candidates"] == len(cands)
    assert stats["keypoints"] == len(kps)
    assert stats["rejSingular"] == 0


def test_reference_default_blur_level(engine):
    """worker.js:35 default min_blur_level = 0.8 (BASELINE configs use 1.6)."""
    ora, cands, kps, _ = _run_case(engine, 96, 80, 3, 42, True, min_blur=0.8)
    check_candidates(cands, ora)
    check_keypoints(kps, ora)


@pytest.mark.parametrize("spo", [1, 2, 4, 5])
def test_other_scales_per_octave(engine, spo):
    """scalesPerOctave != 3: spo + 3 levels per octave, k = 2^(1/spo) (background.js:100, 157), pre-filter
    0.8 * 0.015 * (2^(1/spo) - 1) / (2^(1/3) - 1) (sift.js:285-293); spo = 5 takes the non-TMA scan."""
    u8 = fixtures.synthetic_u8(320, 240, 21)          # enough keypoints for the 99.5 % bar to allow one ill-conditioned walk
    img = fixtures.to_float(u8)
    ora = oracle.detect(img, numberOfOctaves=3, scalesPerOctave=spo, minBlurLevel=1.6, separable=True)
    assert ora.outcomes["singular"] == 0
    prm = L.default_params(numberOfOctaves=3, scalesPerOctave=spo, minBlurLevel=1.6)
    engine.build_scale_space(u8, prm)
    assert engine.pyramid_info() == (3, spo + 3)
    check_levels(engine, ora, L)
    cands, _ = engine.find_candidates()
    thr = 0.8 * 0.015 * (2 ** (1 / spo) - 1) / (2 ** (1 / 3) - 1)
    check_candidates(cands, ora, pix_thr=thr)
    kps, stats = engine.detect(u8, prm)
    check_keypoints(kps, ora, pix_thr=thr)
    assert stats["candidates"] == len(cands) and len(ora.candidates) > 10


def test_f32_and_pitched_inputs(engine):
    """SIFT_F32 pixels, and rows further apart than their length (pitch_bytes): same result as the dense u8 frame."""
    import ctypes as C
    w, h = 150, 90
    u8 = fixtures.synthetic_u8(w, h, 17)
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    want, _ = engine.detect(u8, prm)
    f32, _ = engine.detect((u8.astype(np.float64) / 255.0).astype(np.float32), prm)
    # float32(v / 255) is not v / 255: positions agree to the float32 input rounding, the cells are the same
    assert len(f32) == len(want) and np.array_equal(f32["localX"], want["localX"]) and np.array_equal(f32["localY"], want["localY"])
    assert np.abs(f32["absoluteX"] - want["absoluteX"]).max() < 1e-2
    padded = np.zeros((h, w + 37), np.uint8)
    padded[:, :w] = u8
    out = np.zeros(len(want) + 8, dtype=L.KEYPOINT_DTYPE)
    n = C.c_int()
    st = L.Stats()
    rc = engine._lib.sift_detect(engine._h, padded.ctypes.data, L.SIFT_U8, w, h, padded.strides[0], C.byref(prm),
                                 out.ctypes.data, len(out), C.byref(n), C.byref(st))
    assert rc == L.SIFT_OK and n.value == len(want)
    assert out[:n.value].tobytes() == want.tobytes()


def test_f64_input_matches_u8(engine):
    ora, cands, kps, _ = _run_case(engine, 96, 80, 3, 43, True, dtype="f64")
    check_candidates(cands, ora)
    check_keypoints(kps, ora)


def test_constant_image_has_no_extrema(engine):
    """KAT 1: constant image -> every level equals the constant, ties are never extrema (sift.js:261,266)."""
    img = np.full((40, 56), 0.37, np.float64)
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    engine.build_scale_space(img, prm)
    for o in range(3):
        for s in range(6):
            g = engine.get_level(L.SIFT_LEVEL_GAUSSIAN, o, s)
            assert np.abs(g.astype(np.float64) - 0.37).max() <= 0.37 * 1e-6
    kps, stats = engine.detect(img, prm)
    assert len(kps) == 0 and stats["candidates"] == 0


def test_impulse_support_is_kernel_radius(engine):
    """KAT 2: unit impulse -> outer product of the normalised 1D kernel, support (2R+1)^2, R = round(3 sigma)."""
    img = np.zeros((64, 64), np.float64)
    img[32, 32] = 1.0
    prm = L.default_params(numberOfOctaves=1, minBlurLevel=1.6)
    engine.build_scale_space(img, prm)
    ora = oracle.detect(img, numberOfOctaves=1, minBlurLevel=1.6)
    for s in range(6):
        g = engine.get_level(L.SIFT_LEVEL_GAUSSIAN, 0, s).astype(np.float64)
        R = oracle.kernel_radius(ora.offset_sigma[0][s])
        nz = np.argwhere(g != 0)
        # the impulse is a 2x2 block at (64..65, 64..65) after the nearest-neighbour upsample
        assert nz[:, 0].min() == 64 - R and nz[:, 0].max() == 65 + R
        assert nz[:, 1].min() == 64 - R and nz[:, 1].max() == 65 + R
        assert np.abs(g - ora.gauss[0][s]).max() <= 1e-7


def test_narrow_image_clamps(engine):
    """KAT 3: image narrower than the kernel radius: clamp-to-edge per axis (sift.js:116-119)."""
    u8 = fixtures.synthetic_u8(5, 37, 3)
    img = fixtures.to_float(u8)
    prm = L.default_params(numberOfOctaves=2, minBlurLevel=1.6)
    engine.build_scale_space(u8, prm)
    ora = oracle.detect(img, numberOfOctaves=2, minBlurLevel=1.6)
    check_levels(engine, ora, L)


def test_one_pixel_image(engine):
    img = np.array([[0.25]])
    prm = L.default_params(numberOfOctaves=2, minBlurLevel=1.6)
    kps, stats = engine.detect(img, prm)
    assert len(kps) == 0
    assert engine.octave_size(0) == (2, 2) and engine.octave_size(1) == (1, 1)


def test_detect_is_deterministic_and_ordered(engine):
    u8 = fixtures.synthetic_u8(320, 240, 11)
    prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    a, _ = engine.detect(u8, prm)
    b, _ = engine.detect(u8, prm)
    assert len(a) > 50
    assert a.tobytes() == b.tobytes()


def test_batch_equals_single(engine):
    frames = np.stack([fixtures.synthetic_u8(160, 120, 1234 + i) for i in range(3)])
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    kps, offs, _ = engine.detect_batch(frames, prm)
    for i in range(3):
        single, _ = engine.detect(frames[i], prm)
        assert kps[offs[i]:offs[i + 1]].tobytes() == single.tobytes()


def test_capacity_overflow_is_reported(engine):
    u8 = fixtures.synthetic_u8(320, 240, 11)
    prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    full, _ = engine.detect(u8, prm)
    with pytest.raises(sift_b200.SiftError) as e:
        engine.detect(u8, prm, capacity=3)
    assert e.value.status == L.SIFT_ERR_CAPACITY


def test_bad_arguments(engine):
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=0.4)   # below assumedBlur: sqrt of a negative
    with pytest.raises(sift_b200.SiftError) as e:
        engine.detect(np.zeros((16, 16)), prm)
    assert e.value.status == L.SIFT_ERR_BAD_ARGS
    with pytest.raises(sift_b200.SiftError):
        engine.detect(np.zeros((16, 16)), L.default_params(numberOfOctaves=0))


def test_full_hd_properties(engine):
    """BASELINE configs[1] at full size: size-independent properties (the oracle is too slow here).
    Horizontal mirror symmetry: mirrored input -> mirrored keypoints (the kernels are symmetric and the
    2x upsample / decimation commute with a flip when the octave widths are even)."""
    w, h = 1920, 1080
    u8 = fixtures.synthetic_u8(w, h, 1234)
    prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    a, sa = engine.detect(u8, prm)
    b, sb = engine.detect(np.ascontiguousarray(u8[::-1, :]), prm)   # vertical flip: h = 1080 -> 2160,1080,540,270 all even
    assert len(a) > 1000
    # only octave 0 is flip-symmetric: the decimation in[2a][2b] (matrix2d.js:129) keeps EVEN rows, which a
    # flip of an even-height image maps to odd rows
    ka = {(int(k["scaleLevel"]), int(k["localX"]), int(k["localY"])) for k in a if k["octave"] == 0}
    kb = {(int(k["scaleLevel"]), int(k["localX"]), 2160 - 1 - int(k["localY"])) for k in b if k["octave"] == 0}
    assert len(ka) > 500
    assert len(ka & kb) >= 0.995 * max(len(ka), len(kb))


def test_full_hd_pyramid_is_affine_in_the_image(engine):
    """BASELINE configs[1] at full size, every level of every octave, no oracle needed: the kernels are normalised
    (sift.js:48-63) and everything up to the DoG is linear, so for y = a*x + b every Gaussian level obeys
    G(y) = a*G(x) + b and every DoG level D(y) = a*D(x) -- to the 1e-5 of the level contract."""
    w, h = 1920, 1080
    x = fixtures.to_float(fixtures.synthetic_u8(w, h, 1234))
    a_, b_ = 0.5, 0.25
    y = a_ * x + b_
    prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    worst_g = worst_d = 0.0
    for o in range(4):                                   # one octave at a time keeps the host copies small
        engine.build_scale_space(x, prm)
        gx = [engine.get_level(L.SIFT_LEVEL_GAUSSIAN, o, s).astype(np.float64) for s in range(6)]
        dx = [engine.get_level(L.SIFT_LEVEL_DOG, o, s).astype(np.float64) for s in range(5)]
        engine.build_scale_space(y, prm)
        for s in range(6):
            gy = engine.get_level(L.SIFT_LEVEL_GAUSSIAN, o, s).astype(np.float64)
            want = a_ * gx[s] + b_
            worst_g = max(worst_g, float((np.abs(gy - want) / np.maximum(np.abs(want), 1e-3)).max()))
        for s in range(5):
            dy = engine.get_level(L.SIFT_LEVEL_DOG, o, s).astype(np.float64)
            want = a_ * dx[s]
            worst_d = max(worst_d, float((np.abs(dy - want) / np.maximum(np.abs(want), 0.012)).max()))
    assert worst_g <= 1e-5 and worst_d <= 1e-5, (worst_g, worst_d)
    assert worst_g > 0.0                                 # (the two runs are not trivially the same numbers)


def test_device_resident_frames_in_flight_equal_single(engine):
    """sift_detect_device deals frames to lanes (frames in flight); results must equal the host call, per frame."""
    import torch
    n, w, h = 7, 200, 144
    frames = np.stack([fixtures.synthetic_u8(w, h, 50 + i) for i in range(n)])
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    d = torch.from_numpy(frames).cuda()
    cap = 4096
    out = torch.zeros(n, cap * L.KEYPOINT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(n, dtype=torch.int32, device="cuda")
    for lanes in (3, 1, 4, 8, 0):                                # 0 = chosen by the engine from the frame size
        engine.set_lanes(lanes)
        out.zero_(); cnt.zero_()
        torch.cuda.synchronize()
        for i in range(n):
            engine.detect_device(d[i].data_ptr(), L.SIFT_U8, w, h, 0, prm, out[i].data_ptr(), cap, cnt[i].data_ptr())
        engine.synchronize()
        for i in range(n):
            k = int(cnt[i].item())
            got = np.frombuffer(out[i].cpu().numpy().tobytes(), dtype=L.KEYPOINT_DTYPE)[:k]
            want, _ = engine.detect(frames[i], prm)
            key = lambda a: np.lexsort((a["candX"], a["candY"], a["candScale"], a["octave"]))
            assert k == len(want) and k > 5
            assert got[key(got)].tobytes() == want.tobytes()
    engine.set_lanes(0)


def test_device_resident_ordered_output(engine):
    """sift_detect_device(ordered=1): the device-side sort + gather must give exactly sift_detect's record order
    (octave, scale, row, column: background.js:468-471 / sift.js:221-222), also when cap truncates the output."""
    import torch
    w, h = 320, 240
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    rec = L.KEYPOINT_DTYPE.itemsize
    for seed, cap in ((61, 4096), (62, 4096), (61, 37)):
        u8 = fixtures.synthetic_u8(w, h, seed, blobs=300)
        d = torch.from_numpy(u8).cuda()
        out = torch.zeros(cap * rec, dtype=torch.uint8, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
        torch.cuda.synchronize()
        engine.detect_device(d.data_ptr(), L.SIFT_U8, w, h, 0, prm, out.data_ptr(), cap, cnt.data_ptr(), ordered=True)
        engine.synchronize()
        want, _ = engine.detect(u8, prm)
        k = int(cnt.item())
        got = np.frombuffer(out.cpu().numpy().tobytes(), dtype=L.KEYPOINT_DTYPE)
        if cap >= len(want):
            assert k == len(want) and k > 40
            assert got[:k].tobytes() == want.tobytes()
        else:
            # overflow is reported on the device, never silently truncated: the count is -(capacity that would have
            # sufficed); the cap records that were written are still genuine keypoints, in order
            assert k < 0 and -k >= len(want)
            keys = [(int(r["octave"]), int(r["candScale"]), int(r["candY"]), int(r["candX"])) for r in got[:cap]]
            assert keys == sorted(keys)
            wk = {(int(r["octave"]), int(r["candScale"]), int(r["candY"]), int(r["candX"])) for r in want}
            assert set(keys) <= wk


def test_device_resident_overflow_is_reported_on_the_device(engine):
    """sift_detect_device must not truncate silently.  A `cap` smaller than the number of keypoints (or a
    candidate list larger than the lane's buffer) sets *d_count to -(capacity that would have sufficed)."""
    import torch
    w, h = 320, 240
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=1.6)
    u8 = fixtures.synthetic_u8(w, h, 61, blobs=300)
    want, st = engine.detect(u8, prm)
    d = torch.from_numpy(u8).cuda()
    rec = L.KEYPOINT_DTYPE.itemsize
    for ordered in (False, True):
        for cap in (len(want), len(want) - 1, 5):
            out = torch.zeros(max(cap, 1) * rec, dtype=torch.uint8, device="cuda")
            cnt = torch.zeros(1, dtype=torch.int32, device="cuda")
            torch.cuda.synchronize()
            engine.detect_device(d.data_ptr(), L.SIFT_U8, w, h, 0, prm, out.data_ptr(), cap, cnt.data_ptr(), ordered=ordered)
            engine.synchronize()
            k = int(cnt.item())
            if cap >= len(want):
                assert k == len(want)
            else:
                assert k == -max(st["candidates"], len(want))


def test_stage_replies_follow_the_pyramid_generation(engine):
    """`ss = computeGaussianScaleSpace(A); detect(B); computeDifferenceOfGaussians(ss)` must give A's DoG.
    The reply remembers sift_pyramid_serial(); any call that rebuilds the context's pyramid changes it, and the next
    stage then works from the payload (background.js:258, 359, 455 consume the request) instead of the device."""
    from sift_b200 import background as bg
    a = fixtures.synthetic_u8(96, 80, 5)
    b = fixtures.synthetic_u8(96, 80, 6)
    s0 = engine.pyramid_serial
    ss = bg.computeGaussianScaleSpace(a, 3, 3, 1.6, 0.5, engine=engine)
    assert engine.pyramid_serial != s0 and ss.serial == engine.pyramid_serial
    dog_direct = bg.computeDifferenceOfGaussians(ss, engine=engine)
    cands_direct = bg.findCandidateKeypoints(dog_direct, None, 3, engine=engine)
    kp_direct = bg.refineCandidateKeypoints(dog_direct, 3, 3, cands_direct, 1.6, engine=engine)
    # another image through the same engine, by three different doors
    for other in (lambda: engine.detect(b, L.default_params(numberOfOctaves=3, minBlurLevel=1.6)),
                  lambda: engine.detect_batch(np.stack([b, b]), L.default_params(numberOfOctaves=3, minBlurLevel=1.6)),
                  lambda: bg.computeGaussianScaleSpace(b, 3, 3, 1.6, 0.5, engine=engine)):
        before = engine.pyramid_serial
        other()
        assert engine.pyramid_serial != before
        dog = bg.computeDifferenceOfGaussians(ss, engine=engine)          # from the payload: fp32 levels subtracted
        for o in range(3):
            for s in range(5):
                want = ss[o][s]["image"].astype(np.float64) - ss[o][s + 1]["image"].astype(np.float64)
                assert np.array_equal(np.asarray(dog[o][s]["image"], dtype=np.float64), want)
        other()
        cands = bg.findCandidateKeypoints(dog_direct, None, 3, engine=engine)   # A's DoG is uploaded again
        assert cands == cands_direct
        other()
        kps = bg.refineCandidateKeypoints(dog_direct, 3, 3, cands_direct, 1.6, engine=engine)
        assert kps == kp_direct and len(kps) > 10


def test_batch_overflow_falls_back_and_stays_correct(engine):
    """More keypoints than the device buffers were sized for: the batch path regrows and still matches."""
    frames = np.stack([fixtures.synthetic_u8(96, 96, 300 + i, blobs=400, sigma_lo=0.7, sigma_hi=2.0) for i in range(5)])
    prm = L.default_params(numberOfOctaves=3, minBlurLevel=0.8)
    kps, offs, st = engine.detect_batch(frames, prm)
    assert st["keypoints"] == len(kps) == offs[-1]
    for i in range(5):
        single, _ = engine.detect(frames[i], prm)
        assert kps[offs[i]:offs[i + 1]].tobytes() == single.tobytes()


def test_4k_six_octaves_properties(engine):
    """BASELINE configs[3] at full size (3840x2160, 6 octaves: radii up to 463, wider than the last octave's
    image -> the large-radius fallback kernels run).  Size-independent properties: determinism, reference
    order, every record inside its octave, counters add up, and the octave-0 result does not depend on how
    many octaves follow (octaves only feed forward, background.js:114-130)."""
    w, h = 3840, 2160
    u8 = fixtures.synthetic_u8(w, h, 99)
    p6 = L.default_params(numberOfOctaves=6, minBlurLevel=1.6)
    a, sa = engine.detect(u8, p6)
    b, _ = engine.detect(u8, p6)
    assert a.tobytes() == b.tobytes() and len(a) > 5000
    key = [(int(k["octave"]), int(k["candScale"]), int(k["candY"]), int(k["candX"])) for k in a]
    assert key == sorted(key)
    sizes = [engine.octave_size(o) for o in range(6)]
    assert sizes == [(7680, 4320), (3840, 2160), (1920, 1080), (960, 540), (480, 270), (240, 135)]
    for o, (ow, oh) in enumerate(sizes):
        k = a[a["octave"] == o]
        assert ((k["localX"] >= 1) & (k["localX"] <= ow - 2) & (k["localY"] >= 1) & (k["localY"] <= oh - 2)).all()
        assert ((k["scaleLevel"] >= 1) & (k["scaleLevel"] <= 3)).all()
    assert sa["keypoints"] == len(a)
    assert sa["candidates"] == sa["keypoints"] + sum(sa[r] for r in ("rejLowContrast", "rejEdge", "rejLeftScale", "rejLeftRows",
                                                                   "rejLeftCols", "rejNoConvergence", "rejSingular"))
    assert sa["rejSingular"] == 0
    p2 = L.default_params(numberOfOctaves=2, minBlurLevel=1.6)
    c, _ = engine.detect(u8, p2)
    assert a[a["octave"] <= 1].tobytes() == c.tobytes()


def test_720p_batch_properties(engine):
    """BASELINE configs[2] frame shape (1280x720), a slice of the batch: frames in flight must not change results
    (batch == one frame at a time), per-frame offsets are consistent, distinct seeds give distinct keypoints."""
    n = 6
    frames = np.stack([fixtures.synthetic_u8(1280, 720, 1234 + i) for i in range(n)])
    prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    kps, offs, st = engine.detect_batch(frames, prm)
    assert offs[0] == 0 and offs[-1] == len(kps) == st["keypoints"] and (np.diff(offs) > 500).all()
    for i in (0, 3, 5):
        single, _ = engine.detect(frames[i], prm)
        assert kps[offs[i]:offs[i + 1]].tobytes() == single.tobytes()
    assert kps[offs[0]:offs[1]].tobytes() != kps[offs[1]:offs[2]].tobytes()


def test_keypoint_only_mode_gives_the_same_keypoints(engine):
    """sift_set_keep_gaussian(0): the Gaussian levels stay in registers; DoG, seeds and keypoints are unchanged."""
    u8 = fixtures.synthetic_u8(320, 240, 21)
    prm = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    a, _ = engine.detect(u8, prm)
    engine.set_keep_gaussian(False)
    try:
        b, _ = engine.detect(u8, prm)
    finally:
        engine.set_keep_gaussian(True)
    assert a.tobytes() == b.tobytes() and len(a) > 50
