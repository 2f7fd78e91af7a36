#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by running the UNMODIFIED reference.

Runs /root/reference/{background.js, src/*.js} (read where they lie, never copied) through
oracle/jsmini.py, driving the worker exactly as the page does:

    main.js:98   ImageUtils_convertImageDataToMatrix2D(ImageData)        (rgba fixtures)
    main.js:111  workerComputeGaussianScaleSpace  -> RECEIVED_GAUSSIAN_SCALE_SPACE
    main.js:239  workerComputeDifferenceOfGaussians -> RECEIVED_DIFFERENCE_OF_GAUSSIANS
    main.js:274  workerFindCandidateKeypoints     -> RECEIVED_CANDIDATE_KEYPOINTS
    main.js:325  workerRefineCandidateKeypoints   -> RECEIVED_REFINED_KEYPOINTS

through the reference's own src/worker.js senders and background.js `onmessage` switch.  The worker
globals the browser would provide are shimmed: postMessage (collector), console.log (stub),
OffscreenCanvas(w,h).getContext('2d').createImageData(w,h) (image-utils.js:179-181).

/root/reference only exists in the build container; the committed .npz files are what travels.
Usage:  python oracle/make_golden.py [--ref /root/reference] [--out tests/golden] [--only NAME]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import jsmini  # noqa: E402

# name -> (width, height, seed, blobs, octaves, spo, minBlurLevel, assumedBlur, ingest)
CASES = {   # tiny inputs (the interpreter runs ~60k dense-kernel taps per second); blobs of sigma 0.6-2.2 px
    "g22x18_o2_b16": (22, 18, 15, 30, 2, 3, 1.6, 0.5, "matrix"),      # BASELINE sigma0 = 1.6
    "g28x22_o2_b08": (28, 22, 30, 30, 2, 3, 0.8, 0.5, "matrix"),      # worker.js:35 default sigma0 = 0.8
    "g17x13_o3_b08_rgba": (17, 13, 23, 30, 3, 3, 0.8, 0.5, "rgba"),   # odd sizes (ceil halving) + ImageData ingest
    "g16x16_o2_s2_b10": (16, 16, 31, 30, 2, 2, 1.0, 0.5, "matrix"),   # scalesPerOctave = 2 (threshold formula)
    "g16x14_o2_s4_b10": (16, 14, 73, 30, 2, 4, 1.0, 0.5, "matrix"),   # scalesPerOctave = 4: 7 levels, 6 DoG, scales 1..4
    "g20x16_o2_s1_b12": (20, 16, 37, 30, 2, 1, 1.2, 0.5, "matrix"),   # scalesPerOctave = 1: one tested scale per octave
    "g30x24_o3_b16": (30, 24, 37, 30, 3, 3, 1.6, 0.5, "matrix"),      # BASELINE parameters over 3 octaves: radii up to 58
    # ^ on a 15 x 12 octave (every tap clamps, sift.js:116-119)
    "g40x32_o4_b16": (40, 32, 21, 40, 4, 3, 1.6, 0.5, "matrix"),      # BASELINE configs' full parameter set: 4 octaves, radii
}                                                                     # up to 116 (a 233 x 233 kernel on a 10 x 8 octave)


class _Canvas:
    """OffscreenCanvas shim: only createImageData is used (image-utils.js:179-181)."""

    def __init__(self, w, h):
        self.width, self.height = w, h

    def getContext(self, kind):
        return self

    def createImageData(self, w, h):
        o = jsmini.JSObject()
        o["width"], o["height"] = int(w), int(h)
        o["data"] = jsmini.Uint8ClampedArray(int(w) * int(h) * 4)
        return o


class ReferenceWorker:
    """background.js loaded in a worker-like global scope."""

    def __init__(self, ref_root: str):
        self.outbox = []
        self.log_lines = 0
        console = jsmini.JSObject(log=self._log)
        self.interp = jsmini.Interpreter(ref_root, {"onmessage": None, "postMessage": self._post, "console": console,
                                                    "OffscreenCanvas": _Canvas})
        self.interp.load_module("background.js")
        self.senders = self.interp.load_module("src/worker.js")
        self.image_utils = self.interp.load_module("src/image-utils.js")
        self.sift = self.interp.load_module("src/sift.js")
        self.matrix2d = self.interp.load_module("src/matrix2d.js")
        self.types = self.senders["WorkerMessageTypes"]
        self.handle = jsmini.JSObject(postMessage=self._to_worker)      # what main.js holds as `background_thread`

    def _log(self, *a):
        self.log_lines += 1

    def _post(self, msg, *rest):
        self.outbox.append(msg)

    def _to_worker(self, msg, *rest):
        onmessage = self.interp.get_global("onmessage")
        onmessage(jsmini.JSObject(data=msg))

    def take(self, type_name: str):
        want = self.types[type_name]
        hits = [m for m in self.outbox if m.get("type") == want]
        assert len(hits) == 1, f"{len(hits)} replies of type {want}"
        counts = {}
        for m in self.outbox:
            counts[m["type"]] = counts.get(m["type"], 0) + 1
        self.outbox = []
        return hits[0], counts


def run_reference(ref_root, u8, n_oct, spo, min_blur, assumed, ingest="matrix", chunk=32, verbose=False):
    """Returns dict with every stage output of the reference on the grey image `u8` (H x W uint8)."""
    w = ReferenceWorker(ref_root)
    h_, w_ = u8.shape
    t0 = time.time()
    if ingest == "rgba":
        data = jsmini.Uint8ClampedArray(w_ * h_ * 4)
        rgba = np.repeat(u8.reshape(-1, 1), 4, axis=1).astype(np.uint8)
        rgba[:, 3] = 255
        data.buf = bytearray(rgba.tobytes())
        image_data = jsmini.JSObject(width=w_, height=h_, data=data)
        arg = jsmini.JSObject(imageData=image_data, convertToGrayscale=True, usePerceptualGrayscale=True,
                              discardAlphaChannel=True)
        matrix = w.image_utils["ImageUtils_convertImageDataToMatrix2D"](arg)               # main.js:98-103
    else:
        matrix = [[int(v) / 255.0 for v in row] for row in u8]                              # image-utils.js:114 on a grey byte
    out = {"input_u8": u8.copy(), "input_matrix": np.array(matrix, dtype=np.float64),
           "params": np.array([n_oct, spo, min_blur, assumed], dtype=np.float64), "ingest": ingest}

    w.senders["workerComputeGaussianScaleSpace"](w.handle, jsmini.JSObject(                 # main.js:111-117
        input_image=matrix, min_blur_level=min_blur, chunk_size=chunk, number_of_octaves=n_oct,
        scales_per_octave=spo, assumed_blur=assumed))
    msg, counts = w.take("RECEIVED_GAUSSIAN_SCALE_SPACE")
    scale_space = msg["scaleSpace"]
    out["messages_scale_space"] = counts
    if verbose:
        print(f"  scale space {time.time() - t0:.1f}s", flush=True)

    w.senders["workerComputeDifferenceOfGaussians"](w.handle, scale_space)                  # main.js:239
    msg, counts = w.take("RECEIVED_DIFFERENCE_OF_GAUSSIANS")
    dog = msg["differenceOfGaussians"]
    out["messages_dog"] = counts

    base_images = [octave[0]["image"] for octave in scale_space]                            # main.js:277
    w.senders["workerFindCandidateKeypoints"](w.handle, dog, base_images, spo)              # main.js:274-280
    msg, counts = w.take("RECEIVED_CANDIDATE_KEYPOINTS")
    cands = msg["candidateKeypoints"]
    out["messages_candidates"] = counts

    w.senders["workerRefineCandidateKeypoints"](w.handle, dog, cands, spo, n_oct, min_blur)  # main.js:325-332
    msg, counts = w.take("RECEIVED_REFINED_KEYPOINTS")
    refined = msg["refinedKeypoints"]
    if verbose:
        print(f"  all stages {time.time() - t0:.1f}s, {w.log_lines} console.log lines", flush=True)

    for o, octave in enumerate(scale_space):
        for s, lvl in enumerate(octave):
            out[f"gauss_{o}_{s}"] = np.array(lvl["image"], dtype=np.float64)
            out[f"gauss_blur_{o}_{s}"] = float(lvl["blurLevel"])
    for o, octave in enumerate(dog):
        for s, lvl in enumerate(octave):
            out[f"dog_{o}_{s}"] = np.array(lvl["image"], dtype=np.float64)
            out[f"dog_blur_{o}_{s}"] = float(lvl["blurLevel"])
    rows = []
    for o, octave in enumerate(cands):
        for grp in octave:
            for e in grp["localExtremas"]:
                rows.append((o, grp["scaleLevel"], e["x"], e["y"], e["value"]))
    out["candidates"] = np.array(rows, dtype=np.float64).reshape(-1, 5)
    out["n_low_contrast_markers"] = out["messages_candidates"].get(w.types["RECEIVED_CANDIDATE_KEYPOINT_MARKER"], 0) - len(rows)
    out["keypoints"] = np.array([(k["octave"], k["scaleLevel"], k["localX"], k["localY"], k["absoluteSigma"],
                                  k["absoluteX"], k["absoluteY"], k["interpolatedValue"]) for k in refined],
                                dtype=np.float64).reshape(-1, 8)
    out["n_octaves"] = len(scale_space)
    out["n_levels"] = len(scale_space[0])
    return out


def step_function_vectors(ref_root):
    """Direct calls of the five src/sift.js exports + matrix2d helpers on small inputs."""
    w = ReferenceWorker(ref_root)
    rng = np.random.default_rng(2024)
    out = {}
    img = rng.random((11, 9))
    m = img.tolist()
    output = [[0.0] * 9 for _ in range(11)]
    chunk = w.sift["SIFT_blurMatrix2DChunk"](m, output, 1.3, jsmini.JSObject(x1=2, y1=1, x2=8, y2=10))
    out["blur_in"] = img
    out["blur_sigma"] = 1.3
    out["blur_chunk"] = np.array(chunk)
    out["blur_output"] = np.array(output)
    a, b = rng.random((6, 7)), rng.random((6, 7))
    output = [[0.0] * 7 for _ in range(6)]
    chunk = w.sift["SIFT_subtractMatrix2DChunk"]([a.tolist(), b.tolist()], output, jsmini.JSObject(x1=1, y1=0, x2=7, y2=5))
    out["sub_a"], out["sub_b"], out["sub_chunk"], out["sub_output"] = a, b, np.array(chunk), np.array(output)
    trio = (rng.random((3, 9, 10)) - 0.5) * 0.1
    trio[1, 4, 4] = 0.09
    trio[1, 6, 2] = -0.08
    trio[1, 2, 7] = 0.056
    trio[0, 2, 7] = trio[2, 2, 7] = 0.0
    trio[1, 1:4, 6:9] = np.minimum(trio[1, 1:4, 6:9], 0.05)
    trio[1, 2, 7] = 0.0505
    res = w.sift["SIFT_findExtremas"]([t.tolist() for t in trio], 3)
    out["ext_trio"] = trio
    out["ext_cand"] = np.array([(e["x"], e["y"], e["value"]) for e in res["candidateKeypoints"]]).reshape(-1, 3)
    out["ext_low"] = np.array([(e["x"], e["y"], e["value"]) for e in res["lowContrastKeypoints"]]).reshape(-1, 3)
    dog = [[jsmini.JSObject(image=t.tolist()) for t in trio]]
    out["grad"] = np.array(w.sift["SIFT_generateGradientVector"](0, 1, 4, 5, dog))
    out["hess"] = np.array(w.sift["SIFT_generateHessianMatrix"](0, 1, 4, 5, dog))
    h = rng.random((3, 3)) - 0.5
    h = h + h.T
    out["inv_in"] = h
    out["inv_out"] = np.array(w.matrix2d["Matrix2D_get3x3Inverse"](h.tolist()))
    out["inv_singular_is_null"] = w.matrix2d["Matrix2D_get3x3Inverse"]([[1, 2, 3], [2, 4, 6], [1, 1, 1]]) is None
    r = rng.random((5, 7))
    out["resize_in"] = r
    out["resize_half"] = np.array(w.matrix2d["Matrix2D_linearResize"](r.tolist(), 0.5))
    out["resize_two"] = np.array(w.matrix2d["Matrix2D_linearResize"](r.tolist(), 2.0))
    bounds = w.image_utils["ImageUtils_generateChunkBoundaries"](70, 45, 32)
    out["chunk_bounds_70x45"] = np.array([(c["x1"], c["y1"], c["x2"], c["y2"]) for c in bounds])
    return out


def preview_vectors(ref_root):
    """The display products: Matrix2D_sigmoidNormalize / Matrix2D_sampledNormalize /
    ImageUtils_convertMatrix2DToImageData of the unmodified reference on small matrices."""
    w = ReferenceWorker(ref_root)
    rng = np.random.default_rng(77)
    out = {}
    g = rng.random((9, 13))
    g[0, :6] = [0.0, 1.0, 0.5, 127.5 / 255, 1.2, -0.3]            # exact tie of Math.round, values outside [0, 1]
    d = (rng.random((9, 13)) - 0.5) * 0.3
    conv = w.image_utils["ImageUtils_convertMatrix2DToImageData"]

    def pixels(image_data):
        return np.frombuffer(bytes(image_data["data"].buf), dtype=np.uint8).reshape(9, 13, 4).copy()

    out["gray_in"] = g
    out["gray_rgba"] = pixels(conv(13, 9, jsmini.JSObject(grayChannelMatrix=g.tolist())))
    out["dog_in"] = d
    sig = w.matrix2d["Matrix2D_sigmoidNormalize"](d.tolist(), 5)                      # background.js:303-307
    out["sigmoid_matrix"] = np.array(sig)
    out["sigmoid_rgba"] = pixels(conv(13, 9, jsmini.JSObject(grayChannelMatrix=sig)))
    mm = w.matrix2d["Matrix2D_sampledNormalize"](d.tolist())                          # background.js:336
    out["minmax_matrix"] = np.array(mm)
    out["minmax_rgba"] = pixels(conv(13, 9, jsmini.JSObject(grayChannelMatrix=mm)))
    return out


def chunking_vectors(ref_root):
    """The reference blurs in chunk_size x chunk_size tiles for its progressive repaint (background.js:147-203) but
    reads the whole base image for every tile's halo (sift.js:109-125): run it with two chunk sizes on the same
    input and keep both results, so a test can hold the claim that chunk_size does not change any output."""
    sys.path.insert(0, ROOT)
    from sift_b200 import fixtures
    u8 = fixtures.synthetic_u8(14, 12, 15, blobs=30, sigma_lo=0.6, sigma_hi=2.2)
    out = {"input_u8": u8}
    for chunk in (32, 5):
        res = run_reference(ref_root, u8, 2, 3, 1.0, 0.5, "matrix", chunk=chunk)
        out[f"c{chunk}_keypoints"] = res["keypoints"]
        out[f"c{chunk}_candidates"] = res["candidates"]
        out[f"c{chunk}_levels"] = np.stack([res[f"gauss_0_{s}"] for s in range(6)])
        out[f"c{chunk}_dog1"] = np.stack([res[f"dog_1_{s}"] for s in range(5)])
        out[f"c{chunk}_chunk_messages"] = int(res["messages_scale_space"].get("received-gaussian-blurred-chunk", 0))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    from importlib import import_module
    sys.path.insert(0, ROOT)
    import_module("sift_b200")
    from sift_b200 import fixtures
    os.makedirs(args.out, exist_ok=True)
    if args.only in (None, "steps"):
        t0 = time.time()
        sv = step_function_vectors(args.ref)
        np.savez_compressed(os.path.join(args.out, "ref_steps.npz"), **sv)
        print(f"ref_steps.npz {time.time() - t0:.1f}s", flush=True)
    if args.only in (None, "chunking"):
        np.savez_compressed(os.path.join(args.out, "ref_chunking.npz"), **chunking_vectors(args.ref))
        print("ref_chunking.npz", flush=True)
    if args.only in (None, "preview"):
        np.savez_compressed(os.path.join(args.out, "ref_preview.npz"), **preview_vectors(args.ref))
        print("ref_preview.npz", flush=True)
    for name, (w_, h_, seed, blobs, n_oct, spo, mb, ab, ingest) in CASES.items():
        if args.only not in (None, name):
            continue
        t0 = time.time()
        u8 = fixtures.synthetic_u8(w_, h_, seed, blobs=blobs, sigma_lo=0.6, sigma_hi=2.2)
        res = run_reference(args.ref, u8, n_oct, spo, mb, ab, ingest, verbose=True)
        flat = {k: v for k, v in res.items() if not k.startswith("messages_")}
        for stage in ("scale_space", "dog", "candidates"):
            m = res["messages_" + stage]
            flat["msgcount_" + stage] = np.array(sorted(m.items()), dtype=object).astype(str)
        np.savez_compressed(os.path.join(args.out, f"ref_{name}.npz"), **flat)
        print(f"ref_{name}.npz: {len(res['candidates'])} candidates, {len(res['keypoints'])} keypoints, "
              f"{time.time() - t0:.1f}s", flush=True)


if __name__ == "__main__":
    main()
