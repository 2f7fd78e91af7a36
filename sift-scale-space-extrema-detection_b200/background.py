"""Host mirror of the reference's pipeline driver background.js: the four stage
functions behind its onmessage switch (background.js:14-50), with the same option
names and reply schemas, plus the fused `detect` that keeps everything on the GPU.
The postMessage transport (src/worker.js) is replaced by direct calls.

  computeGaussianScaleSpace     background.js:71
  computeDifferenceOfGaussians  background.js:258
  findCandidateKeypoints        background.js:359
  refineCandidateKeypoints      background.js:455
"""
from __future__ import annotations

import numpy as np

from . import _lib as L
from .engine import Engine, default_engine


class _Resident(list):
    """A reply (nested lists, like the reference's) that also remembers which engine built it and the engine's
    pyramid generation at that moment (sift_pyramid_serial), so the next stage does not re-upload a pyramid the
    device still holds -- and DOES re-upload it when anything (another image, detect(), a strip) replaced it."""
    engine: Engine | None = None
    params: L.Params | None = None
    serial: int = -1


def _params(number_of_octaves=5, scales_per_octave=3, min_blur_level=0.8, assumed_blur=0.5, **extra) -> L.Params:
    # defaults: worker.js:33-37
    return L.default_params(numberOfOctaves=int(number_of_octaves), scalesPerOctave=int(scales_per_octave),
                            minBlurLevel=float(min_blur_level), assumedBlur=float(assumed_blur), **extra)


# message types of the per-level display products (src/worker.js:9, 14)
RECEIVED_GAUSSIAN_BLURRED_IMAGE = "received-gaussian-blurred-image"
RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE = "received-difference-of-gaussian-image"


def _post_level_images(eng: Engine, kind: int, post_message):
    """The ImageData the reference posts per level: Gaussian levels as grey (background.js:
    136-143, 212-220), DoG levels min/max-normalised (background.js:331-338); {type, imageData, octave}."""
    n_oct, nlev = eng.pyramid_info()
    dog = kind == L.SIFT_LEVEL_DOG
    for o in range(n_oct):
        for s in range(nlev - (1 if dog else 0)):
            rgba, _ = eng.level_preview(kind, o, s, L.SIFT_PREVIEW_MINMAX if dog else L.SIFT_PREVIEW_GRAY)
            post_message({"type": RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE if dog else RECEIVED_GAUSSIAN_BLURRED_IMAGE,
                          "imageData": {"width": rgba.shape[1], "height": rgba.shape[0], "data": rgba},
                          "octave": o})


def dogChunkPreview(octave: int, scale: int, chunk: dict, engine: Engine | None = None):
    """One chunk of a resident DoG level as the reference paints it while subtracting: sigmoid-normalised,
    coefficient 5 (background.js:303-317; matrix2d.js:148-156).  chunk = {x1, y1, x2, y2}, half-open."""
    eng = engine or default_engine()
    rgba, _ = eng.level_preview(L.SIFT_LEVEL_DOG, octave, scale, L.SIFT_PREVIEW_SIGMOID, 5.0)
    crop = np.ascontiguousarray(rgba[chunk["y1"]:chunk["y2"], chunk["x1"]:chunk["x2"]])
    return {"imageData": {"width": crop.shape[1], "height": crop.shape[0], "data": crop}, "dx": chunk["x1"], "dy": chunk["y1"]}


def computeGaussianScaleSpace(input_image, number_of_octaves=5, scales_per_octave=3, min_blur_level=0.8,
                              assumed_blur=0.5, chunk_size=32, engine: Engine | None = None, post_message=None,
                              **thresholds):
    """scale_space[o][s] = {blurLevel, image} (background.js:57-70, 233-236).  chunk_size only
    shaped the reference's progressive repaint (background.js:181-203); results do not depend on it.
    post_message: optional callable receiving the per-level RECEIVED_GAUSSIAN_BLURRED_IMAGE messages."""
    eng = engine or default_engine()
    prm = _params(number_of_octaves, scales_per_octave, min_blur_level, assumed_blur, **thresholds)
    eng.build_scale_space(input_image, prm)
    n_oct, nlev = eng.pyramid_info()
    reply = _Resident()
    for o in range(n_oct):
        reply.append([{"blurLevel": eng.blur_level(L.SIFT_LEVEL_GAUSSIAN, o, s),
                       "image": eng.get_level(L.SIFT_LEVEL_GAUSSIAN, o, s)} for s in range(nlev)])
    reply.engine, reply.params, reply.serial = eng, prm, eng.pyramid_serial
    if post_message is not None:
        _post_level_images(eng, L.SIFT_LEVEL_GAUSSIAN, post_message)
    return reply


def _is_resident(obj, eng) -> bool:
    return isinstance(obj, _Resident) and obj.engine is eng and obj.serial >= 0 and eng.pyramid_serial == obj.serial


def computeDifferenceOfGaussians(scale_space, chunk_size=32, engine: Engine | None = None, post_message=None):
    """D[o][s-1] = S[o][s-1] - S[o][s], blurLevel of S[o][s-1] (background.js:258-354).
    post_message: optional callable receiving the per-level RECEIVED_DIFFERENCE_OF_GAUSSIAN_IMAGE messages
    (resident pyramids only)."""
    eng = engine or (scale_space.engine if isinstance(scale_space, _Resident) and scale_space.engine else default_engine())
    reply = _Resident()
    if _is_resident(scale_space, eng):
        # the blur kernels already formed the DoG from the unrounded accumulators
        eng.build_dog()
        n_oct, nlev = eng.pyramid_info()
        for o in range(n_oct):
            reply.append([{"blurLevel": eng.blur_level(L.SIFT_LEVEL_DOG, o, s),
                           "image": eng.get_level(L.SIFT_LEVEL_DOG, o, s)} for s in range(nlev - 1)])
        reply.engine, reply.params, reply.serial = eng, scale_space.params, scale_space.serial
        if post_message is not None:
            _post_level_images(eng, L.SIFT_LEVEL_DOG, post_message)
        return reply
    # foreign scale space: subtract level pairs on the GPU (SIFT_subtractMatrix2DChunk over the whole image)
    for octave in scale_space:
        levels = []
        for s in range(1, len(octave)):
            a = np.ascontiguousarray(np.asarray(octave[s - 1]["image"], dtype=np.float64))
            b = np.ascontiguousarray(np.asarray(octave[s]["image"], dtype=np.float64))
            out = np.zeros_like(a)
            eng.subtract_chunk(a, b, out, 0, 0, a.shape[1], a.shape[0])
            levels.append({"blurLevel": octave[s - 1]["blurLevel"], "image": out})
        reply.append(levels)
    return reply


def _ensure_dog_resident(difference_of_gaussians, eng: Engine, scales_per_octave: int, min_blur_level=0.8):
    """Make `eng` hold these DoG levels (upload when they are not the engine's own)."""
    if _is_resident(difference_of_gaussians, eng):
        return difference_of_gaussians.params
    n_oct = len(difference_of_gaussians)
    h0, w0 = np.asarray(difference_of_gaussians[0][0]["image"]).shape
    prm = _params(n_oct, scales_per_octave, min_blur_level, 0.5)
    eng.set_pyramid_shape(w0, h0, prm)
    for o, octave in enumerate(difference_of_gaussians):
        for s, lvl in enumerate(octave):
            eng.set_level(L.SIFT_LEVEL_DOG, o, s, lvl["image"])
    if isinstance(difference_of_gaussians, _Resident):
        # adopted: the device now holds exactly these levels (fp32, as they were read back)
        difference_of_gaussians.engine, difference_of_gaussians.params = eng, prm
        difference_of_gaussians.serial = eng.pyramid_serial
    return prm


def findCandidateKeypoints(differenceOfGaussians, octaveBaseImages=None, scalesPerOctave=3,
                           engine: Engine | None = None, want_low_contrast=False):
    """candidateKeypoints[o][i] = {scaleLevel, localExtremas:[{x,y,value}]} (background.js:359-450).
    octaveBaseImages is accepted and ignored, as in the reference (background.js:361)."""
    eng = engine or (differenceOfGaussians.engine if isinstance(differenceOfGaussians, _Resident) and
                     differenceOfGaussians.engine else default_engine())
    _ensure_dog_resident(differenceOfGaussians, eng, scalesPerOctave)
    cands, low = eng.find_candidates(want_low_contrast=want_low_contrast)
    n_oct = len(differenceOfGaussians)
    n_scales = len(differenceOfGaussians[0])
    reply = [[{"scaleLevel": s, "localExtremas": []} for s in range(1, n_scales - 1)] for _ in range(n_oct)]
    for c in cands:
        reply[int(c["octave"])][int(c["scaleLevel"]) - 1]["localExtremas"].append(
            {"x": int(c["x"]), "y": int(c["y"]), "value": float(c["value"])})
    if want_low_contrast:
        lows = [[{"scaleLevel": s, "localExtremas": []} for s in range(1, n_scales - 1)] for _ in range(n_oct)]
        for c in low:
            lows[int(c["octave"])][int(c["scaleLevel"]) - 1]["localExtremas"].append(
                {"x": int(c["x"]), "y": int(c["y"]), "value": float(c["value"])})
        return reply, lows
    return reply


def keypoint_records(kps: np.ndarray) -> list:
    """Structured array -> the reference's record dicts (background.js:619-628) + the extra fields."""
    out = []
    for k in kps:
        out.append({"octave": int(k["octave"]), "scaleLevel": int(k["scaleLevel"]), "localX": int(k["localX"]),
                    "localY": int(k["localY"]), "absoluteSigma": float(k["absoluteSigma"]),
                    "absoluteX": float(k["absoluteX"]), "absoluteY": float(k["absoluteY"]),
                    "interpolatedValue": float(k["interpolatedValue"]),
                    "offset": [float(v) for v in k["offset"]], "dogValue": float(k["dogValue"])})
    return out


def refineCandidateKeypoints(differenceOfGaussians, scalesPerOctave, numberOfOctaves, candidateKeypoints,
                             minBlurLevel, minInterpixelDistance=0.5, engine: Engine | None = None,
                             return_stats=False):
    """refinedKeypoints[] (background.js:455-685)."""
    eng = engine or (differenceOfGaussians.engine if isinstance(differenceOfGaussians, _Resident) and
                     differenceOfGaussians.engine else default_engine())
    _ensure_dog_resident(differenceOfGaussians, eng, scalesPerOctave, minBlurLevel)
    flat = []
    for octave in range(int(numberOfOctaves)):                         # background.js:468-471
        for scale_i in range(int(scalesPerOctave)):
            entry = candidateKeypoints[octave][scale_i]
            for e in entry["localExtremas"]:
                flat.append((octave, int(entry["scaleLevel"]), int(e["x"]), int(e["y"]), float(e["value"]), 0))
    cands = np.array(flat, dtype=L.CANDIDATE_DTYPE) if flat else np.zeros(0, dtype=L.CANDIDATE_DTYPE)
    # thresholds ride in the context's params; positions need minBlurLevel / minInterpixelDistance
    prm = L.default_params(numberOfOctaves=int(numberOfOctaves), scalesPerOctave=int(scalesPerOctave),
                           minBlurLevel=float(minBlurLevel), minInterpixelDistance=float(minInterpixelDistance))
    kps, stats = eng.refine(cands, prm)
    recs = keypoint_records(kps)
    return (recs, stats) if return_stats else recs


def detect(input_image, number_of_octaves=5, scales_per_octave=3, min_blur_level=0.8, assumed_blur=0.5,
           engine: Engine | None = None, as_records=True, rgba=False, **thresholds):
    """The whole chain main.js:111 -> 239 -> 274 -> 325 in one device-resident call."""
    eng = engine or default_engine()
    prm = _params(number_of_octaves, scales_per_octave, min_blur_level, assumed_blur, **thresholds)
    kps, stats = eng.detect(input_image, prm, rgba=rgba)
    return (keypoint_records(kps) if as_records else kps), stats
