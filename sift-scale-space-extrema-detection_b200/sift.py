"""Host mirror of the reference's src/sift.js exports (same names, argument meaning
and in-place behaviour), forwarding to the CUDA engine through the C ABI.

  SIFT_blurMatrix2DChunk        src/sift.js:72
  SIFT_subtractMatrix2DChunk    src/sift.js:154
  SIFT_findExtremas             src/sift.js:212
  SIFT_generateGradientVector   src/sift.js:333
  SIFT_generateHessianMatrix    src/sift.js:377

A Matrix2D is a list of rows (matrix2d.js:5-31) or a 2-D numpy array.  Lists are the
compat (slow) form; float64 ndarrays avoid the per-element conversion.
"""
from __future__ import annotations

import numpy as np

from .engine import default_engine


def _is_list(m) -> bool:
    return isinstance(m, list)


def _chunk(cb):
    return int(cb["x1"]), int(cb["y1"]), int(cb["x2"]), int(cb["y2"])


def _write_back(output, buf: np.ndarray, x1, y1, x2, y2):
    """`output[y][x] = value` for the chunk (sift.js:137, :173) on the caller's own object."""
    if isinstance(output, np.ndarray):
        if output is not buf:
            output[y1:y2, x1:x2] = buf[y1:y2, x1:x2]
    else:
        for y in range(y1, y2):
            output[y][x1:x2] = buf[y, x1:x2].tolist()


def SIFT_blurMatrix2DChunk(input, output, sigma, chunk_boundary, engine=None):
    """Blur the half-open chunk of `input` with a Gaussian of `sigma` (radius round(3 sigma),
    clamp-to-edge); mutates `output` and returns the chunk matrix (sift.js:72-149)."""
    eng = engine or default_engine()
    x1, y1, x2, y2 = _chunk(chunk_boundary)
    a = np.ascontiguousarray(np.asarray(input, dtype=np.float64))
    direct = isinstance(output, np.ndarray) and output.dtype == np.float64 and output.flags.c_contiguous
    buf = output if direct else np.zeros_like(a)
    eng.blur_chunk(a, buf, sigma, x1, y1, x2, y2)
    _write_back(output, buf, x1, y1, x2, y2)
    chunk = buf[y1:y2, x1:x2]
    return chunk.tolist() if _is_list(input) else chunk.copy()


def SIFT_subtractMatrix2DChunk(input_pair, output, chunk_boundary, engine=None):
    """output = input_pair[0] - input_pair[1] over the chunk (sift.js:154-188)."""
    eng = engine or default_engine()
    x1, y1, x2, y2 = _chunk(chunk_boundary)
    a = np.ascontiguousarray(np.asarray(input_pair[0], dtype=np.float64))
    b = np.ascontiguousarray(np.asarray(input_pair[1], dtype=np.float64))
    direct = isinstance(output, np.ndarray) and output.dtype == np.float64 and output.flags.c_contiguous
    buf = output if direct else np.zeros_like(a)
    eng.subtract_chunk(a, b, buf, x1, y1, x2, y2)
    _write_back(output, buf, x1, y1, x2, y2)
    chunk = buf[y1:y2, x1:x2]
    return chunk.tolist() if _is_list(input_pair[0]) else chunk.copy()


def SIFT_findExtremas(image_trio, scales_per_octave, engine=None):
    """Strict 26-neighbour extrema of image_trio[1], raster order, split by the 0.8 * threshold
    pre-filter (sift.js:212-316). Returns {candidateKeypoints, lowContrastKeypoints}."""
    eng = engine or default_engine()
    (cxy, cv), (lxy, lv) = eng.find_extremas(image_trio[0], image_trio[1], image_trio[2], int(scales_per_octave))
    mk = lambda xy, v: [{"x": int(xy[i, 0]), "y": int(xy[i, 1]), "value": float(v[i])} for i in range(len(v))]
    return {"candidateKeypoints": mk(cxy, cv), "lowContrastKeypoints": mk(lxy, lv)}


def _dog_image(difference_of_gaussians, o, s):
    lvl = difference_of_gaussians[o][s]
    return lvl["image"] if isinstance(lvl, dict) else lvl.image


def SIFT_generateGradientVector(o, s, m, n, difference_of_gaussians, engine=None):
    """[ds, dm, dn] central differences (sift.js:333-353)."""
    eng = engine or default_engine()
    g, _ = eng.gradient_hessian(_dog_image(difference_of_gaussians, o, s - 1), _dog_image(difference_of_gaussians, o, s),
                                _dog_image(difference_of_gaussians, o, s + 1), int(m), int(n))
    return g.tolist()


def SIFT_generateHessianMatrix(o, s, m, n, difference_of_gaussians, engine=None):
    """Symmetric 3x3 finite-difference Hessian in [s, m, n] order (sift.js:377-447)."""
    eng = engine or default_engine()
    _, h = eng.gradient_hessian(_dog_image(difference_of_gaussians, o, s - 1), _dog_image(difference_of_gaussians, o, s),
                                _dog_image(difference_of_gaussians, o, s + 1), int(m), int(n))
    return h.tolist()
