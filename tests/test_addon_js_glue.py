"""addon/*.js executed.  Node.js is absent from this image, so the ES modules a Node host would import -- native.js,
sift.js, background.js, worker-adapter.js -- run under oracle/jsmini.py (the interpreter that runs the reference) with
`require('./sift_b200.node')` answered by a Python stand-in for the N-API addon:
  * stage calls answer from the reference's own outputs (tests/golden/ref_g22x18_o2_b16.npz), as fp32 planes and packed
    24 / 80-byte records exactly as sift_addon.c hands them over, so the replies the glue assembles can be compared
    with the replies the reference produced;
  * step calls are computed by the float64 oracle, so SIFT_blurMatrix2DChunk & co. can be compared with the reference's
    step vectors (tests/golden/ref_steps.npz).
What this covers is the glue: record decoding / encoding, typed-array plumbing, in-place semantics, reply schemas and
the worker message protocol.  The arithmetic behind the addon is covered by the GPU tests."""
import os

import numpy as np
import pytest

import oracle
from oracle import jsmini
from sift_b200 import _lib as L

import js_host as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADDON = os.path.join(ROOT, "addon")
GOLDEN = os.path.join(ROOT, "tests", "golden")
J = jsmini.JSObject


class FakeAddon:
    """Same entry points and value shapes as addon/sift_addon.c."""

    def __init__(self, g):
        self.g = g
        self.n_oct, self.nlev = int(g["n_octaves"]), int(g["n_levels"])
        self.calls = []
        self.serial = 0                    # sift_pyramid_serial: every build / detect / setLevel bumps it
        self.uploaded = {}                 # (kind, o, s) -> plane set through setLevel

    # -- context / fused
    def create(self, device=0):
        self.calls.append("create")
        return J(handle=1)

    def _records(self, rows):
        out = np.zeros(len(rows), dtype=L.KEYPOINT_DTYPE)
        for i, r in enumerate(rows):
            out[i]["octave"], out[i]["scaleLevel"], out[i]["localX"], out[i]["localY"] = (int(v) for v in r[:4])
            out[i]["absoluteSigma"], out[i]["absoluteX"], out[i]["absoluteY"], out[i]["interpolatedValue"] = r[4:8]
            out[i]["offset"] = (0.25, -0.5, 0.125)
            out[i]["dogValue"] = 0.0625
        return H.ArrayBuffer(np.frombuffer(out.tobytes(), dtype=np.uint8).copy())

    def detect(self, ctx, data, w, h, dtype, prm):
        self.calls.append(("detect", w, h, dtype, prm["numberOfOctaves"], prm["minBlurLevel"]))
        self.serial += 1
        k = self.g["keypoints"]
        return J(records=self._records(k), count=len(k), stats=J(keypoints=len(k), candidates=len(self.g["candidates"])))

    def detectBatch(self, ctx, frames, w, h, n_images, dtype, prm, capacity):
        assert frames.length == w * h * n_images
        k = self.g["keypoints"]
        counts = [len(k) if i % 2 == 0 else 0 for i in range(int(n_images))]          # odd frames: no keypoints
        rows = np.concatenate([k] * sum(1 for c in counts if c)) if any(counts) else k[:0]
        offs = H.Int32Array.of_numpy(np.concatenate([[0], np.cumsum(counts)]).astype(np.int32))
        return J(records=self._records(rows), offsets=offs, stats=J(keypoints=len(rows)))

    # -- stages
    def buildScaleSpace(self, ctx, data, w, h, dtype, prm):
        assert isinstance(data, H.Float64Array) and dtype == 2                       # a Matrix2D arrives as F64
        assert np.array_equal(data.a.reshape(h, w), self.g["input_matrix"])
        assert (prm["numberOfOctaves"], prm["scalesPerOctave"], prm["minBlurLevel"], prm["assumedBlur"]) == tuple(self.g["params"])
        assert prm["contrastThreshold"] == 0.015 and prm["edgeRatio"] == 10 and prm["maxIterations"] == 5
        self.calls.append("buildScaleSpace")
        self.serial += 1

    def pyramidSerial(self, ctx):
        return self.serial

    def setPyramidShape(self, ctx, w0, h0, prm):
        self.calls.append(("setPyramidShape", int(w0), int(h0), int(prm["numberOfOctaves"]), int(prm["scalesPerOctave"])))
        self.serial += 1

    def setLevel(self, ctx, kind, o, s, data):
        assert isinstance(data, H.Float32Array)
        self.uploaded[(int(kind), int(o), int(s))] = data.a.copy()
        self.serial += 1

    def pyramidInfo(self, ctx):
        return J(octaves=self.n_oct, levels=self.nlev)

    def _level(self, kind, o, s):
        return self.g[("gauss_%d_%d" if kind == 0 else "dog_%d_%d") % (o, s)]

    def getLevel(self, ctx, kind, o, s):
        m = self._level(kind, o, s)
        blur = float(self.g[("gauss_blur_%d_%d" if kind == 0 else "dog_blur_%d_%d") % (o, s)])
        return J(blurLevel=blur, width=m.shape[1], height=m.shape[0], data=H.Float32Array.of_numpy(m.astype(np.float32).ravel()))

    def levelPreview(self, ctx, kind, o, s, mode, coefficient):
        rgba, mm = oracle.preview(self._level(kind, o, s), int(mode), float(coefficient))
        return J(width=rgba.shape[1], height=rgba.shape[0], min=mm[0], max=mm[1], data=H.Uint8ClampedArray.of_numpy(rgba.ravel()))

    def findCandidates(self, ctx, prm, want_low):
        c = self.g["candidates"]
        rec = np.zeros(len(c), dtype=L.CANDIDATE_DTYPE)
        rec["octave"], rec["scaleLevel"], rec["x"], rec["y"], rec["value"] = c[:, 0], c[:, 1], c[:, 2], c[:, 3], c[:, 4]
        return J(count=len(c), records=H.ArrayBuffer(np.frombuffer(rec.tobytes(), dtype=np.uint8).copy()))

    def refine(self, ctx, prm, buffer, count):
        got = np.frombuffer(buffer.data.tobytes(), dtype=L.CANDIDATE_DTYPE)[:int(count)]    # what encodeCandidates packed
        c = self.g["candidates"]
        assert int(count) == len(c)
        assert np.array_equal(np.stack([got["octave"], got["scaleLevel"], got["x"], got["y"]], axis=1), c[:, :4].astype(int))
        assert np.array_equal(got["value"], c[:, 4].astype(np.float32))
        self.calls.append("refine")
        k = self.g["keypoints"]
        return J(count=len(k), records=self._records(k))

    # -- steps, computed by the oracle on the Float64Arrays the glue passes
    def blurChunk(self, ctx, src, rows, cols, dst, sigma, x1, y1, x2, y2):
        out = dst.a.reshape(rows, cols).copy()
        oracle.blur_chunk(src.a.reshape(rows, cols), out, sigma, x1, y1, x2, y2)
        dst.a[:] = out.ravel()

    def subtractChunk(self, ctx, a, b, rows, cols, dst, x1, y1, x2, y2):
        out = dst.a.reshape(rows, cols).copy()
        oracle.subtract_chunk(a.a.reshape(rows, cols), b.a.reshape(rows, cols), out, x1, y1, x2, y2)
        dst.a[:] = out.ravel()

    def findExtremas(self, ctx, d0, d1, d2, rows, cols, spo, contrast, prefactor):
        r = oracle.find_extremas([d.a.reshape(rows, cols) for d in (d0, d1, d2)], int(spo), contrast, prefactor)
        cand, low = r["candidateKeypoints"], r["lowContrastKeypoints"]

        def pack(lst):
            return (H.Int32Array.of_numpy(np.array([(e["x"], e["y"]) for e in lst], dtype=np.int32).ravel()),
                    H.Float64Array.of_numpy(np.array([e["value"] for e in lst], dtype=np.float64)))
        cxy, cv = pack(cand)
        lxy, lv = pack(low)
        return J(nCand=len(cand), candXY=cxy, candValue=cv, nLow=len(low), lowXY=lxy, lowValue=lv)

    def gradientHessian(self, ctx, dm, dc, dp, rows, cols, m, n):
        trio = [d.a.reshape(rows, cols) for d in (dm, dc, dp)]
        g, h = oracle.gradient(trio, 1, int(m), int(n)), oracle.hessian(trio, 1, int(m), int(n))
        return H.Float64Array.of_numpy(np.concatenate([g, h.ravel()]))


def test_stage_requests_are_consumed_not_the_contexts_last_pyramid(golden):
    """the reference consumes the pyramid IN each request (background.js:258, 359, 455).  The glue may
    skip the upload only for the reply it produced itself while nothing rebuilt the context's pyramid; after an
    intervening detect() of another image the scale space is re-subtracted from the payload and the DoG uploaded."""
    fake = FakeAddon(golden)
    interp, _ = H.make_interpreter(ADDON, fake)
    bg = interp.load_module("background.js")
    n_oct, spo, min_blur, assumed = (float(v) for v in golden["params"])
    req = J(inputImage=[list(map(float, r)) for r in golden["input_matrix"]], numberOfOctaves=n_oct, scalesPerOctave=spo,
            minBlurLevel=min_blur, assumedBlur=assumed, chunkSize=32.0)
    ss = bg["computeGaussianScaleSpace"](req)
    dog = bg["computeDifferenceOfGaussians"](ss)                      # resident: read back, no subtraction call
    bg["findCandidateKeypoints"](J(differenceOfGaussians=dog, scalesPerOctave=spo))
    assert not fake.uploaded and not any(isinstance(c, tuple) and c[0] == "setPyramidShape" for c in fake.calls)
    # another image goes through the context
    bg["detect"](J(data=H.Uint8Array(16 * 12), width=16.0, height=12.0), J(numberOfOctaves=2.0))
    dog2 = bg["computeDifferenceOfGaussians"](ss)                     # recomputed from the payload, pair by pair
    for o in range(fake.n_oct):
        for s in range(fake.nlev - 1):
            want = golden["gauss_%d_%d" % (o, s)].astype(np.float32).astype(np.float64) - \
                golden["gauss_%d_%d" % (o, s + 1)].astype(np.float32).astype(np.float64)
            assert np.array_equal(np.array(dog2[o][s]["image"]), want)
            assert dog2[o][s]["blurLevel"] == ss[o][s]["blurLevel"]
    bg["findCandidateKeypoints"](J(differenceOfGaussians=dog, scalesPerOctave=spo))     # uploaded: all DoG levels
    shape = [c for c in fake.calls if isinstance(c, tuple) and c[0] == "setPyramidShape"]
    h0, w0 = golden["dog_0_0"].shape
    assert shape == [("setPyramidShape", w0, h0, fake.n_oct, int(spo))]
    assert set(fake.uploaded) == {(1, o, s) for o in range(fake.n_oct) for s in range(fake.nlev - 1)}
    assert np.array_equal(fake.uploaded[(1, 0, 1)], golden["dog_0_1"].astype(np.float32).ravel())
    n_up = len(fake.uploaded); fake.uploaded.clear()
    bg["refineCandidateKeypoints"](J(differenceOfGaussians=dog, scalesPerOctave=spo, numberOfOctaves=n_oct,
                                     candidateKeypoints=bg["findCandidateKeypoints"](J(differenceOfGaussians=dog, scalesPerOctave=spo)),
                                     minBlurLevel=min_blur))
    assert not fake.uploaded and n_up > 0                            # the uploaded pyramid is now the resident one


@pytest.fixture()
def golden():
    return np.load(os.path.join(GOLDEN, "ref_g22x18_o2_b16.npz"), allow_pickle=True)


def test_worker_adapter_speaks_the_reference_protocol(golden):
    fake = FakeAddon(golden)
    interp, drain = H.make_interpreter(ADDON, fake)
    mod = interp.load_module("worker-adapter.js")
    T = mod["WorkerMessageTypes"]
    n_oct, spo, min_blur, assumed = (float(v) for v in golden["params"])
    inbox = []
    worker = mod["SiftWorker"](J(previews=True))
    worker["onmessage"] = lambda e: inbox.append(e["data"])

    def post(**m):
        worker["postMessage"](J(**m))
        assert inbox == [] or inbox[-1]["type"] != m["type"]      # replies are asynchronous, like a Worker's
        drain()
        out = list(inbox)
        inbox.clear()
        return out

    # main.js:111-117 -> background.js:71
    msgs = post(type=T["COMPUTE_GAUSSIAN_SCALE_SPACE"], inputImage=golden["input_matrix"].tolist(), numberOfOctaves=n_oct,
                scalesPerOctave=spo, minBlurLevel=min_blur, assumedBlur=assumed, chunkSize=32)
    assert [m["type"] for m in msgs] == [T["RECEIVED_GAUSSIAN_BLURRED_IMAGE"]] * 12 + [T["RECEIVED_GAUSSIAN_SCALE_SPACE"]]
    ss = msgs[-1]["scaleSpace"]
    assert len(ss) == 2 and all(len(o) == 6 for o in ss)
    for o in range(2):
        for s in range(6):
            ref = golden[f"gauss_{o}_{s}"]
            assert ss[o][s]["blurLevel"] == float(golden[f"gauss_blur_{o}_{s}"])
            assert np.array_equal(np.array(ss[o][s]["image"]), ref.astype(np.float32).astype(np.float64))   # Matrix2D rows
    img = msgs[0]["imageData"]
    want, _ = oracle.preview(golden["gauss_0_0"], oracle.PREVIEW_GRAY)
    assert (img["width"], img["height"]) == (44, 36) and msgs[0]["octave"] == 0 and msgs[11]["octave"] == 1
    assert isinstance(img["data"], H.Uint8ClampedArray) and np.array_equal(img["data"].a, want.ravel())

    # main.js:239 -> background.js:258
    msgs = post(type=T["COMPUTE_DIFFERENCE_OF_GAUSSIANS"], scaleSpace=ss)
    assert [m["type"] for m in msgs] == [T["RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE"]] * 10 + [T["RECEIVED_DIFFERENCE_OF_GAUSSIANS"]]
    dog = msgs[-1]["differenceOfGaussians"]
    assert [len(o) for o in dog] == [5, 5]
    assert dog[1][2]["blurLevel"] == float(golden["dog_blur_1_2"])
    want, _ = oracle.preview(golden["dog_1_4"], oracle.PREVIEW_MINMAX)
    assert np.array_equal(msgs[9]["imageData"]["data"].a, want.ravel())

    # main.js:274-280 -> background.js:359: candidateKeypoints[o][i] = {scaleLevel, localExtremas:[{x, y, value}]}
    msgs = post(type=T["FIND_CANDIDATE_KEYPOINTS"], differenceOfGaussians=dog, octaveBaseImages=[], scalesPerOctave=spo)
    cands = msgs[-1]["candidateKeypoints"]
    assert [m["type"] for m in msgs] == [T["RECEIVED_CANDIDATE_KEYPOINTS"]]
    assert [[e["scaleLevel"] for e in o] for o in cands] == [[1, 2, 3], [1, 2, 3]]
    flat = [(o, e["scaleLevel"], x["x"], x["y"], x["value"]) for o, oc in enumerate(cands) for e in oc for x in e["localExtremas"]]
    ref = golden["candidates"]
    assert len(flat) == len(ref) > 0
    assert np.array_equal(np.array(flat)[:, :4], ref[:, :4]) and np.array_equal(np.array(flat)[:, 4], ref[:, 4].astype(np.float32))

    # main.js:325-332 -> background.js:455: the eight fields of background.js:619-628 (+ offset, dogValue)
    msgs = post(type=T["REFINE_CANDIDATE_KEYPOINTS"], differenceOfGaussians=dog, candidateKeypoints=cands, scalesPerOctave=spo,
                numberOfOctaves=n_oct, minBlurLevel=min_blur)
    kps = msgs[-1]["refinedKeypoints"]
    ref = golden["keypoints"]
    assert msgs[-1]["type"] == T["RECEIVED_REFINED_KEYPOINTS"] and len(kps) == len(ref) > 0
    fields = ("octave", "scaleLevel", "localX", "localY", "absoluteSigma", "absoluteX", "absoluteY", "interpolatedValue")
    assert np.array_equal(np.array([[k[f] for f in fields] for k in kps]), ref)
    assert kps[0]["offset"] == [0.25, -0.5, 0.125] and kps[0]["dogValue"] == 0.0625
    assert fake.calls.count("create") == 1 and "refine" in fake.calls          # one context, like the reference's one worker
    worker["postMessage"](J(type="no-such-request"))                              # background.js:18-49: unknown types are ignored
    drain()
    assert inbox == []


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_the_references_own_senders_drive_the_adapter(golden):
    """the reference's src/worker.js sender functions (what main.js calls, worker.js:29-98), unmodified,
    post their requests to a SiftWorker instead of a Worker; the message types of both files agree."""
    fake = FakeAddon(golden)
    interp, drain = H.make_interpreter(ADDON, fake)
    ours = interp.load_module("worker-adapter.js")
    ref = interp.load_module("/root/reference/src/worker.js")
    for name, value in ours["WorkerMessageTypes"].items():
        assert ref["WorkerMessageTypes"][name] == value, name
    n_oct, spo, min_blur, assumed = (float(v) for v in golden["params"])
    inbox = []
    worker = ours["SiftWorker"]()
    worker["onmessage"] = lambda e: inbox.append(e["data"])
    ref["workerComputeGaussianScaleSpace"](worker, J(input_image=golden["input_matrix"].tolist(), min_blur_level=min_blur,
                                                     chunk_size=32, number_of_octaves=n_oct, scales_per_octave=spo,
                                                     assumed_blur=assumed))                        # main.js:111-117
    drain()
    ss = inbox.pop()["scaleSpace"]
    ref["workerComputeDifferenceOfGaussians"](worker, ss)                                          # main.js:239
    drain()
    dog = inbox.pop()["differenceOfGaussians"]
    ref["workerFindCandidateKeypoints"](worker, dog, [o[0]["image"] for o in ss], spo)             # main.js:274-280
    drain()
    cands = inbox.pop()["candidateKeypoints"]
    ref["workerRefineCandidateKeypoints"](worker, dog, cands, spo, n_oct, min_blur)                # main.js:325-332
    drain()
    msg = inbox.pop()
    assert inbox == [] and msg["type"] == ref["WorkerMessageTypes"]["RECEIVED_REFINED_KEYPOINTS"]
    fields = ("octave", "scaleLevel", "localX", "localY", "absoluteSigma", "absoluteX", "absoluteY", "interpolatedValue")
    assert np.array_equal(np.array([[k[f] for f in fields] for k in msg["refinedKeypoints"]]), golden["keypoints"])


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="reference tree only exists in the build container")
def test_option_a_the_references_driver_with_our_step_functions():
    """INTEGRATION.md option A, executed: the reference's UNMODIFIED background.js with its `./src/sift.js` import
    answered by addon/sift.js (arithmetic behind it: the float64 oracle standing in for the addon).  Driven like the
    page drives its worker, the four replies equal the reference's own run on the same input, bit for bit."""
    from oracle import make_golden
    g = np.load(os.path.join(GOLDEN, "ref_g16x16_o2_s2_b10.npz"), allow_pickle=True)
    outbox = []
    interp, _ = H.make_interpreter(ADDON, FakeAddon(g))
    interp.global_scope.vars.update({"onmessage": None, "postMessage": lambda m, *r: outbox.append(m),
                                     "console": J(log=lambda *a: None), "OffscreenCanvas": make_golden._Canvas})
    interp.virtual_modules["./src/sift.js"] = interp.load_module("sift.js")        # the swap of background.js:6
    interp.load_module("/root/reference/background.js")
    send = interp.load_module("/root/reference/src/worker.js")
    T = send["WorkerMessageTypes"]
    handle = J(postMessage=lambda m, *r: interp.get_global("onmessage")(J(data=m)))
    n_oct, spo, min_blur, assumed = (float(v) for v in g["params"])

    def reply(kind):
        hits = [m for m in outbox if m.get("type") == T[kind]]
        outbox.clear()
        assert len(hits) == 1
        return hits[0]

    send["workerComputeGaussianScaleSpace"](handle, J(input_image=g["input_matrix"].tolist(), min_blur_level=min_blur,
                                                      chunk_size=32, number_of_octaves=n_oct, scales_per_octave=spo,
                                                      assumed_blur=assumed))
    ss = reply("RECEIVED_GAUSSIAN_SCALE_SPACE")["scaleSpace"]
    for o in range(int(n_oct)):
        for s in range(int(spo) + 3):
            assert np.array_equal(np.array(ss[o][s]["image"]), g[f"gauss_{o}_{s}"]), (o, s)
    send["workerComputeDifferenceOfGaussians"](handle, ss)
    dog = reply("RECEIVED_DIFFERENCE_OF_GAUSSIANS")["differenceOfGaussians"]
    assert np.array_equal(np.array(dog[1][2]["image"]), g["dog_1_2"])
    send["workerFindCandidateKeypoints"](handle, dog, [o[0]["image"] for o in ss], spo)
    cands = reply("RECEIVED_CANDIDATE_KEYPOINTS")["candidateKeypoints"]
    flat = [(o, e["scaleLevel"], x["x"], x["y"], x["value"]) for o, oc in enumerate(cands) for e in oc for x in e["localExtremas"]]
    assert np.array_equal(np.array(flat), g["candidates"])
    send["workerRefineCandidateKeypoints"](handle, dog, cands, spo, n_oct, min_blur)
    kps = reply("RECEIVED_REFINED_KEYPOINTS")["refinedKeypoints"]
    fields = ("octave", "scaleLevel", "localX", "localY", "absoluteSigma", "absoluteX", "absoluteY", "interpolatedValue")
    got = np.array([[k[f] for f in fields] for k in kps])
    assert got.shape == g["keypoints"].shape and np.array_equal(got[:, :4], g["keypoints"][:, :4])
    assert np.array_equal(got[:, 5:], g["keypoints"][:, 5:]) and np.allclose(got[:, 4], g["keypoints"][:, 4], rtol=4e-16, atol=0)


def test_fused_detect_and_pixel_type_dispatch(golden):
    fake = FakeAddon(golden)
    interp, _ = H.make_interpreter(ADDON, fake)
    bg = interp.load_module("background.js")
    nat = interp.load_module("native.js")
    u8 = H.Uint8Array.of_numpy(golden["input_u8"].ravel())
    r = bg["detect"](J(data=u8, width=22, height=18), J(numberOfOctaves=2, minBlurLevel=1.6))
    assert len(r["keypoints"]) == len(golden["keypoints"]) and r["stats"]["keypoints"] == len(golden["keypoints"])
    assert fake.calls[-1] == ("detect", 22, 18, nat["DTYPE"]["U8"], 2, 1.6)
    frames = H.Uint8Array(22 * 18 * 3)
    b = bg["detectBatch"](frames, 22, 18, 3, nat["DTYPE"]["U8"], J(numberOfOctaves=2))
    nk = len(golden["keypoints"])
    assert [len(f) for f in b["keypoints"]] == [nk, 0, nk] and b["stats"]["keypoints"] == 2 * nk
    assert b["keypoints"][2][0]["absoluteX"] == float(golden["keypoints"][0][5])
    rgba = H.Uint8ClampedArray(22 * 18 * 4)
    assert nat["toPixels"](J(data=rgba, width=22, height=18))["dtype"] == nat["DTYPE"]["RGBA8"]     # ImageData
    assert nat["toPixels"](J(data=H.Float32Array(22 * 18), width=22, height=18))["dtype"] == nat["DTYPE"]["F32"]
    assert nat["toPixels"]([[0.0, 1.0], [0.5, 0.25]])["dtype"] == nat["DTYPE"]["F64"]
    with pytest.raises(Exception):
        nat["toPixels"](J(data=[1, 2, 3], width=3, height=1))


def test_step_functions_keep_the_reference_semantics(golden):
    """src/sift.js:72-149: `output` is mutated in place inside the chunk only and the chunk comes back as a matrix."""
    s = np.load(os.path.join(GOLDEN, "ref_steps.npz"), allow_pickle=True)
    interp, _ = H.make_interpreter(ADDON, FakeAddon(golden))
    sift = interp.load_module("sift.js")
    out = [[0.0] * 9 for _ in range(11)]
    chunk = sift["SIFT_blurMatrix2DChunk"](s["blur_in"].tolist(), out, float(s["blur_sigma"]), J(x1=2, y1=1, x2=8, y2=10))
    assert np.array_equal(np.array(chunk), s["blur_chunk"]) and np.array_equal(np.array(out), s["blur_output"])
    out = [[0.0] * 7 for _ in range(6)]
    chunk = sift["SIFT_subtractMatrix2DChunk"]([s["sub_a"].tolist(), s["sub_b"].tolist()], out, J(x1=1, y1=0, x2=7, y2=5))
    assert np.array_equal(np.array(chunk), s["sub_chunk"]) and np.array_equal(np.array(out), s["sub_output"])
    # the fast shape: {data: Float64Array, width, height} in, written in place, chunk back in the same shape
    src = J(data=H.Float64Array.of_numpy(s["sub_a"].ravel()), width=7, height=6)
    src2 = J(data=H.Float64Array.of_numpy(s["sub_b"].ravel()), width=7, height=6)
    dst = J(data=H.Float64Array(42), width=7, height=6)
    chunk = sift["SIFT_subtractMatrix2DChunk"]([src, src2], dst, J(x1=1, y1=0, x2=7, y2=5))
    assert np.array_equal(dst["data"].a.reshape(6, 7), s["sub_output"]) and (chunk["width"], chunk["height"]) == (6, 5)
    assert np.array_equal(chunk["data"].a.reshape(5, 6), s["sub_chunk"])
    r = sift["SIFT_findExtremas"]([t.tolist() for t in s["ext_trio"]], 3)
    got = np.array([(e["x"], e["y"], e["value"]) for e in r["candidateKeypoints"]]).reshape(-1, 3)
    low = np.array([(e["x"], e["y"], e["value"]) for e in r["lowContrastKeypoints"]]).reshape(-1, 3)
    assert np.array_equal(got, s["ext_cand"]) and np.array_equal(low, s["ext_low"])
    dog = [[J(image=t.tolist()) for t in s["ext_trio"]]]
    assert np.array_equal(np.array(sift["SIFT_generateGradientVector"](0, 1, 4, 5, dog)), s["grad"])
    assert np.array_equal(np.array(sift["SIFT_generateHessianMatrix"](0, 1, 4, 5, dog)), s["hess"])


class EngineAddon:
    """The stage entry points of sift_addon.c over the real engine (ctypes): what a Node host would get from the GPU."""

    def __init__(self, engine):
        self.e = engine

    def create(self, device=0):
        return J(handle=1)

    @staticmethod
    def _params(p):
        return L.default_params(numberOfOctaves=int(p["numberOfOctaves"]), scalesPerOctave=int(p["scalesPerOctave"]),
                                minBlurLevel=float(p["minBlurLevel"]), assumedBlur=float(p["assumedBlur"]))

    def buildScaleSpace(self, ctx, data, w, h, dtype, prm):
        self.prm = self._params(prm)
        self.e.build_scale_space(np.ascontiguousarray(data.a.reshape(int(h), int(w))), self.prm)
        self.e.build_dog()

    def pyramidInfo(self, ctx):
        n_oct, nlev = self.e.pyramid_info()
        return J(octaves=n_oct, levels=nlev)

    def pyramidSerial(self, ctx):
        return self.e.pyramid_serial

    def setPyramidShape(self, ctx, w0, h0, prm):
        self.prm = self._params(prm)
        self.e.set_pyramid_shape(int(w0), int(h0), self.prm)

    def setLevel(self, ctx, kind, o, s, data):
        w, h = self.e.octave_size(int(o))
        self.e.set_level(int(kind), int(o), int(s), data.a.reshape(h, w))

    def subtractChunk(self, ctx, a, b, rows, cols, dst, x1, y1, x2, y2):
        out = dst.a.reshape(int(rows), int(cols))
        self.e.subtract_chunk(np.ascontiguousarray(a.a.reshape(int(rows), int(cols))),
                              np.ascontiguousarray(b.a.reshape(int(rows), int(cols))), out, int(x1), int(y1), int(x2), int(y2))

    def getLevel(self, ctx, kind, o, s):
        m = self.e.get_level(int(kind), int(o), int(s))
        return J(blurLevel=self.e.blur_level(int(kind), int(o), int(s)), width=m.shape[1], height=m.shape[0],
                 data=H.Float32Array.of_numpy(m.ravel()))

    def levelPreview(self, ctx, kind, o, s, mode, coefficient):
        rgba, mm = self.e.level_preview(int(kind), int(o), int(s), int(mode), float(coefficient))
        return J(width=rgba.shape[1], height=rgba.shape[0], min=mm[0], max=mm[1], data=H.Uint8ClampedArray.of_numpy(rgba.ravel()))

    def findCandidates(self, ctx, prm, want_low):
        c, _ = self.e.find_candidates()
        return J(count=len(c), records=H.ArrayBuffer(np.frombuffer(c.tobytes(), dtype=np.uint8).copy()))

    def refine(self, ctx, prm, buffer, count):
        c = np.frombuffer(buffer.data.tobytes(), dtype=L.CANDIDATE_DTYPE)[:int(count)]
        k, _ = self.e.refine(c, self.prm)
        return J(count=len(k), records=H.ArrayBuffer(np.frombuffer(k.tobytes(), dtype=np.uint8).copy()))


@pytest.mark.gpu
def test_worker_protocol_over_the_real_engine_reproduces_the_reference(engine, golden):
    """main.js's four requests -> addon/worker-adapter.js -> addon/background.js -> (stand-in for the N-API layer) ->
    libsift_b200.so on the GPU: the refined keypoints are the reference's (same cells, positions to 1e-3 px)."""
    interp, drain = H.make_interpreter(ADDON, EngineAddon(engine))
    mod = interp.load_module("worker-adapter.js")
    T = mod["WorkerMessageTypes"]
    n_oct, spo, min_blur, assumed = (float(v) for v in golden["params"])
    inbox = []
    worker = mod["SiftWorker"](J(previews=True))
    worker["onmessage"] = lambda e: inbox.append(e["data"])

    def post(**m):
        worker["postMessage"](J(**m))
        drain()
        out = list(inbox)
        inbox.clear()
        return out

    msgs = post(type=T["COMPUTE_GAUSSIAN_SCALE_SPACE"], inputImage=golden["input_matrix"].tolist(), numberOfOctaves=n_oct,
                scalesPerOctave=spo, minBlurLevel=min_blur, assumedBlur=assumed, chunkSize=32)
    assert len(msgs) == 13
    ss = msgs[-1]["scaleSpace"]
    for o in range(2):
        for s in range(6):
            ref = golden[f"gauss_{o}_{s}"]
            assert np.abs(np.array(ss[o][s]["image"]) - ref).max() <= 1e-5 * np.abs(ref).max()
    dog = post(type=T["COMPUTE_DIFFERENCE_OF_GAUSSIANS"], scaleSpace=ss)[-1]["differenceOfGaussians"]
    cands = post(type=T["FIND_CANDIDATE_KEYPOINTS"], differenceOfGaussians=dog, octaveBaseImages=[], scalesPerOctave=spo)[-1]["candidateKeypoints"]
    flat = [(o, e["scaleLevel"], x["x"], x["y"]) for o, oc in enumerate(cands) for e in oc for x in e["localExtremas"]]
    assert np.array_equal(np.array(flat), golden["candidates"][:, :4])
    kps = post(type=T["REFINE_CANDIDATE_KEYPOINTS"], differenceOfGaussians=dog, candidateKeypoints=cands, scalesPerOctave=spo,
               numberOfOctaves=n_oct, minBlurLevel=min_blur)[-1]["refinedKeypoints"]
    ref = golden["keypoints"]
    got = np.array([[k[f] for f in ("octave", "scaleLevel", "localX", "localY", "absoluteSigma", "absoluteX", "absoluteY",
                                    "interpolatedValue")] for k in kps])
    assert got.shape == ref.shape and np.array_equal(got[:, :4], ref[:, :4])
    assert np.abs(got[:, 5:7] - ref[:, 5:7]).max() <= 1e-3 and np.allclose(got[:, 4], ref[:, 4], rtol=1e-3)
