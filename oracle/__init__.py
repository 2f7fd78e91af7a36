"""TEST INFRASTRUCTURE ONLY -- ctypes wrapper over oracle/sift_oracle.c.

The float64 CPU restatement of the reference detection path (background.js,
src/sift.js, src/matrix2d.js).  Only tests/, ``__graft_entry__.smoke()`` and
``bench.py``'s cpu_baseline / ``--impl reference`` legs import this package; the
product package never does (tests/test_no_oracle_in_product.py enforces it).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsift_oracle.so")

MAX_OCTAVES = 12
MAX_LEVELS = 12

OUTCOMES = ("accepted", "low_contrast", "edge", "left_scale", "left_rows", "left_cols",
            "no_convergence", "singular")


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (oracle/Makefile)."""
    src = os.path.join(_HERE, "sift_oracle.c")
    hdr = os.path.join(_HERE, "sift_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libsift_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Extremum(C.Structure):
    _fields_ = [("x", C.c_int), ("y", C.c_int), ("value", C.c_double)]


class Candidate(C.Structure):
    _fields_ = [("octave", C.c_int), ("scale", C.c_int), ("x", C.c_int), ("y", C.c_int),
                ("value", C.c_double)]


class Keypoint(C.Structure):
    _fields_ = [("octave", C.c_int), ("scaleLevel", C.c_int), ("localX", C.c_int), ("localY", C.c_int),
                ("absoluteSigma", C.c_double), ("absoluteX", C.c_double), ("absoluteY", C.c_double),
                ("interpolatedValue", C.c_double), ("offset", C.c_double * 3), ("dogValue", C.c_double),
                ("candScale", C.c_int), ("candX", C.c_int), ("candY", C.c_int), ("iterations", C.c_int)]


class Params(C.Structure):
    _fields_ = [("numberOfOctaves", C.c_int), ("scalesPerOctave", C.c_int),
                ("minBlurLevel", C.c_double), ("assumedBlur", C.c_double),
                ("contrastThreshold", C.c_double), ("preFilterFactor", C.c_double),
                ("edgeRatio", C.c_double), ("maxIterations", C.c_int),
                ("offsetBound", C.c_double), ("minInterpixelDistance", C.c_double)]


_DP = C.POINTER(C.c_double)


class Pyramid(C.Structure):
    _fields_ = [("octaves", C.c_int), ("levels", C.c_int), ("spo", C.c_int),
                ("rows", C.c_int * MAX_OCTAVES), ("cols", C.c_int * MAX_OCTAVES),
                ("gauss", (_DP * MAX_LEVELS) * MAX_OCTAVES),
                ("blur", (C.c_double * MAX_LEVELS) * MAX_OCTAVES),
                ("offset_sigma", (C.c_double * MAX_LEVELS) * MAX_OCTAVES),
                ("dog", (_DP * MAX_LEVELS) * MAX_OCTAVES),
                ("dog_blur", (C.c_double * MAX_LEVELS) * MAX_OCTAVES)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.oracle_js_round.restype = C.c_double
        L.oracle_js_round.argtypes = [C.c_double]
        L.oracle_preview.argtypes = [_DP, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, _DP]
        L.oracle_contrast_threshold.restype = C.c_double
        L.oracle_contrast_threshold.argtypes = [C.c_int, C.c_double]
        L.oracle_kernel_radius.argtypes = [C.c_double]
        L.oracle_build_gaussian_kernel.argtypes = [C.c_double, _DP]
        L.oracle_linear_resize_dims.argtypes = [C.c_int, C.c_int, C.c_double, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_linear_resize.argtypes = [_DP, C.c_int, C.c_int, C.c_double, _DP]
        L.oracle_blur_chunk.argtypes = [_DP, C.c_int, C.c_int, _DP, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_blur_image.argtypes = [_DP, C.c_int, C.c_int, _DP, C.c_double]
        L.oracle_blur_image_separable.argtypes = [_DP, C.c_int, C.c_int, _DP, C.c_double]
        L.oracle_subtract_chunk.argtypes = [_DP, _DP, C.c_int, C.c_int, _DP, C.c_int, C.c_int, C.c_int, C.c_int]
        L.oracle_find_extremas.argtypes = [_DP, _DP, _DP, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                           C.POINTER(Extremum), C.c_int, C.POINTER(C.c_int),
                                           C.POINTER(Extremum), C.c_int, C.POINTER(C.c_int)]
        L.oracle_gradient.restype = None
        L.oracle_gradient.argtypes = [C.POINTER(_DP), C.c_int, C.c_int, C.c_int, C.c_int, _DP]
        L.oracle_hessian.restype = None
        L.oracle_hessian.argtypes = [C.POINTER(_DP), C.c_int, C.c_int, C.c_int, C.c_int, _DP]
        L.oracle_inverse3x3.argtypes = [_DP, _DP]
        L.oracle_pyramid_free.restype = None
        L.oracle_pyramid_free.argtypes = [C.POINTER(Pyramid)]
        L.oracle_compute_gaussian_scale_space.argtypes = [_DP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                                          C.c_double, C.c_int, C.POINTER(Pyramid)]
        L.oracle_compute_dog.argtypes = [C.POINTER(Pyramid)]
        L.oracle_find_candidates.argtypes = [C.POINTER(Pyramid), C.c_double, C.c_double,
                                             C.POINTER(Candidate), C.c_int, C.POINTER(C.c_int),
                                             C.POINTER(Candidate), C.c_int, C.POINTER(C.c_int)]
        L.oracle_refine.argtypes = [C.POINTER(Pyramid), C.POINTER(Candidate), C.c_int,
                                    C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                    C.POINTER(Keypoint), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.oracle_refine_one.argtypes = [C.POINTER(Pyramid), C.POINTER(Candidate),
                                        C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.POINTER(Keypoint)]
        L.oracle_refine_margin.restype = C.c_double
        L.oracle_refine_margin.argtypes = [C.POINTER(Pyramid), C.POINTER(Candidate), C.c_double, C.c_double, C.c_int,
                                           C.c_double]
        L.oracle_detect.argtypes = [_DP, C.c_int, C.c_int, C.POINTER(Params), C.c_int, C.POINTER(Pyramid),
                                    C.POINTER(Keypoint), C.c_int, C.POINTER(C.c_int),
                                    C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        _lib = L
    return _lib


def _dp(a: np.ndarray):
    return a.ctypes.data_as(_DP)


def _img(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    assert a.ndim == 2
    return a


def default_params(**kw) -> Params:
    """worker.js:33-37 defaults + the hard-coded constants (sift.js:285,293; background.js:461,480,558,598)."""
    p = Params(5, 3, 0.8, 0.5, 0.015, 0.8, 10.0, 5, 0.6, 0.5)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(k)
        setattr(p, k, v)
    return p


# ------------------------------------------------------------------ step functions

def js_round(x: float) -> float:
    return lib().oracle_js_round(float(x))


PREVIEW_GRAY, PREVIEW_SIGMOID, PREVIEW_MINMAX = 0, 1, 2


def preview(m, mode: int = PREVIEW_GRAY, coefficient: float = 1.0):
    """RGBA8 ImageData bytes [rows, cols, 4] of a matrix + (min, max) (image-utils.js:171-220, matrix2d.js:148-192)."""
    a = _img(m)
    out = np.zeros(a.shape + (4,), dtype=np.uint8)
    mm = np.zeros(2, dtype=np.float64)
    if lib().oracle_preview(_dp(a), a.shape[0], a.shape[1], mode, float(coefficient), out.ctypes.data, _dp(mm)) != 0:
        raise ValueError("oracle_preview: bad arguments")
    return out, (float(mm[0]), float(mm[1]))


def kernel_radius(sigma: float) -> int:
    return lib().oracle_kernel_radius(float(sigma))


def build_gaussian_kernel(sigma: float) -> np.ndarray:
    r = kernel_radius(sigma)
    k = np.empty((2 * r + 1, 2 * r + 1), np.float64)
    lib().oracle_build_gaussian_kernel(float(sigma), _dp(k))
    return k


def linear_resize(m, rate: float) -> np.ndarray:
    m = _img(m)
    r, c = C.c_int(), C.c_int()
    lib().oracle_linear_resize_dims(m.shape[0], m.shape[1], float(rate), C.byref(r), C.byref(c))
    out = np.empty((r.value, c.value), np.float64)
    lib().oracle_linear_resize(_dp(m), m.shape[0], m.shape[1], float(rate), _dp(out))
    return out


def blur_chunk(inp, out: np.ndarray, sigma: float, x1: int, y1: int, x2: int, y2: int) -> np.ndarray:
    """SIFT_blurMatrix2DChunk: mutates ``out`` and returns the chunk (sift.js:72-149)."""
    inp = _img(inp)
    assert out.dtype == np.float64 and out.flags.c_contiguous and out.shape == inp.shape
    lib().oracle_blur_chunk(_dp(inp), inp.shape[0], inp.shape[1], _dp(out), float(sigma), x1, y1, x2, y2)
    return out[y1:y2, x1:x2].copy()


def blur_image(inp, sigma: float, separable: bool = False) -> np.ndarray:
    inp = _img(inp)
    out = np.zeros_like(inp)
    f = lib().oracle_blur_image_separable if separable else lib().oracle_blur_image
    f(_dp(inp), inp.shape[0], inp.shape[1], _dp(out), float(sigma))
    return out


def subtract_chunk(a, b, out: np.ndarray, x1: int, y1: int, x2: int, y2: int) -> np.ndarray:
    a, b = _img(a), _img(b)
    lib().oracle_subtract_chunk(_dp(a), _dp(b), a.shape[0], a.shape[1], _dp(out), x1, y1, x2, y2)
    return out[y1:y2, x1:x2].copy()


def find_extremas(trio, spo: int = 3, contrast: float = 0.015, prefactor: float = 0.8):
    d = [_img(t) for t in trio]
    rows, cols = d[0].shape
    cap = max(1, rows * cols)
    ec = (Extremum * cap)()
    el = (Extremum * cap)()
    nc, nl = C.c_int(), C.c_int()
    lib().oracle_find_extremas(_dp(d[0]), _dp(d[1]), _dp(d[2]), rows, cols, spo, contrast, prefactor,
                               ec, cap, C.byref(nc), el, cap, C.byref(nl))
    cv = lambda arr, n: [{"x": arr[i].x, "y": arr[i].y, "value": arr[i].value} for i in range(n)]
    return {"candidateKeypoints": cv(ec, nc.value), "lowContrastKeypoints": cv(el, nl.value)}


def _dogptrs(dog_levels):
    keep = [_img(d) for d in dog_levels]
    arr = (_DP * len(keep))(*[_dp(k) for k in keep])
    return keep, arr


def gradient(dog_levels, s: int, m: int, n: int) -> np.ndarray:
    keep, arr = _dogptrs(dog_levels)
    g = np.empty(3)
    lib().oracle_gradient(arr, keep[0].shape[1], s, m, n, _dp(g))
    return g


def hessian(dog_levels, s: int, m: int, n: int) -> np.ndarray:
    keep, arr = _dogptrs(dog_levels)
    h = np.empty((3, 3))
    lib().oracle_hessian(arr, keep[0].shape[1], s, m, n, _dp(h))
    return h


def inverse3x3(m):
    m = np.ascontiguousarray(m, np.float64)
    inv = np.empty((3, 3))
    ok = lib().oracle_inverse3x3(_dp(m), _dp(inv))
    return inv if ok else None


# ------------------------------------------------------------------ pipeline

def _kp_dict(k: Keypoint) -> dict:
    return {"octave": k.octave, "scaleLevel": k.scaleLevel, "localX": k.localX, "localY": k.localY,
            "absoluteSigma": k.absoluteSigma, "absoluteX": k.absoluteX, "absoluteY": k.absoluteY,
            "interpolatedValue": k.interpolatedValue, "offset": [k.offset[0], k.offset[1], k.offset[2]],
            "dogValue": k.dogValue, "candScale": k.candScale, "candX": k.candX, "candY": k.candY,
            "iterations": k.iterations}


def refine_on_dog(dog_levels, octave: int, candidate: dict, params: Params | None = None, **kw):
    """oracle_refine_one (background.js:455-685) on caller-supplied DoG levels of one octave (float64 arrays, all the
    same shape) -- e.g. the float32 levels a device stores, to tell its arithmetic from its storage precision.
    Returns (outcome name, keypoint dict or None)."""
    prm = params if params is not None else default_params(**kw)
    keep = [_img(d) for d in dog_levels]
    pyr = Pyramid()
    pyr.octaves, pyr.levels, pyr.spo = octave + 1, len(keep) + 1, prm.scalesPerOctave
    pyr.rows[octave], pyr.cols[octave] = keep[0].shape
    for s, k in enumerate(keep):
        pyr.dog[octave][s] = _dp(k)
    c = Candidate(octave, int(candidate["scale"] if "scale" in candidate else candidate["scaleLevel"]),
                  int(candidate["x"]), int(candidate["y"]), float(candidate["value"]))
    kp = Keypoint()
    rc = lib().oracle_refine_one(C.byref(pyr), C.byref(c), prm.contrastThreshold, prm.edgeRatio, prm.maxIterations,
                                 prm.offsetBound, prm.minBlurLevel, prm.minInterpixelDistance, C.byref(kp))
    return OUTCOMES[rc], (_kp_dict(kp) if rc == 0 else None)


class Result:
    """Everything the four stages produce for one image.  The level arrays are views of the C pyramid, which is
    released with the Result (close() / garbage collection)."""

    def __init__(self):
        self.gauss = []        # [o][s] float64 arrays
        self.blur = []         # [o][s] blurLevel
        self.offset_sigma = []
        self.dog = []          # [o][s]
        self.candidates = []   # dicts octave, scale, x, y, value (reference order)
        self.n_low_contrast = 0
        self.low_contrast = [] # dicts like candidates: extrema below the pre-filter (sift.js:301-306), reference order
        self.margins = {}      # (octave, scale, y, x) -> oracle_refine_margin of that candidate (test diagnostic)
        self.keypoints = []    # dicts (reference order)
        self.outcomes = {}
        self._pyr = None

    def close(self):
        if self._pyr is not None:
            self.gauss, self.dog = [], []
            lib().oracle_pyramid_free(C.byref(self._pyr))
            self._pyr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def detect(image, params: Params | None = None, separable: bool = False, keep_levels: bool = True,
           **kw) -> Result:
    """Run the whole reference path (main.js:111 -> 239 -> 274 -> 325) on a [0,1] float image."""
    img = _img(image)
    prm = params if params is not None else default_params(**kw)
    L = lib()
    pyr = Pyramid()
    rc = L.oracle_compute_gaussian_scale_space(_dp(img), img.shape[0], img.shape[1], prm.numberOfOctaves,
                                               prm.scalesPerOctave, prm.minBlurLevel, prm.assumedBlur,
                                               1 if separable else 0, C.byref(pyr))
    if rc:
        raise RuntimeError(f"oracle scale space failed rc={rc}")
    res = Result()
    res._pyr = pyr
    try:
        L.oracle_compute_dog(C.byref(pyr))
        nc, nl = C.c_int(), C.c_int()
        L.oracle_find_candidates(C.byref(pyr), prm.contrastThreshold, prm.preFilterFactor,
                                 None, 0, C.byref(nc), None, 0, C.byref(nl))
        cand = (Candidate * max(1, nc.value))()
        low = (Candidate * max(1, nl.value))()
        L.oracle_find_candidates(C.byref(pyr), prm.contrastThreshold, prm.preFilterFactor,
                                 cand, nc.value, C.byref(nc), low, nl.value, C.byref(nl))
        kps = (Keypoint * max(1, nc.value))()
        nk = C.c_int()
        outc = (C.c_int * len(OUTCOMES))()
        L.oracle_refine(C.byref(pyr), cand, nc.value, prm.contrastThreshold, prm.edgeRatio, prm.maxIterations,
                        prm.offsetBound, prm.minBlurLevel, prm.minInterpixelDistance,
                        kps, nc.value, C.byref(nk), outc)
        for o in range(pyr.octaves):
            shape = (pyr.rows[o], pyr.cols[o])
            n = shape[0] * shape[1]
            if keep_levels:
                res.gauss.append([np.ctypeslib.as_array(pyr.gauss[o][s], shape=(n,)).reshape(shape)
                                  for s in range(pyr.levels)])
                res.dog.append([np.ctypeslib.as_array(pyr.dog[o][s], shape=(n,)).reshape(shape)
                                for s in range(pyr.levels - 1)])
            res.blur.append([pyr.blur[o][s] for s in range(pyr.levels)])
            res.offset_sigma.append([pyr.offset_sigma[o][s] for s in range(pyr.levels)])
        res.shapes = [(pyr.rows[o], pyr.cols[o]) for o in range(pyr.octaves)]
        res.candidates = [{"octave": cand[i].octave, "scale": cand[i].scale, "x": cand[i].x, "y": cand[i].y,
                           "value": cand[i].value} for i in range(nc.value)]
        res.n_low_contrast = nl.value
        res.low_contrast = [{"octave": low[i].octave, "scale": low[i].scale, "x": low[i].x, "y": low[i].y,
                             "value": low[i].value} for i in range(nl.value)]
        for i in range(nc.value):
            res.margins[(cand[i].octave, cand[i].scale, cand[i].y, cand[i].x)] = L.oracle_refine_margin(
                C.byref(pyr), C.byref(cand[i]), prm.contrastThreshold, prm.edgeRatio, prm.maxIterations,
                prm.offsetBound)
        res.keypoints = [_kp_dict(kps[i]) for i in range(nk.value)]
        res.outcomes = {OUTCOMES[i]: outc[i] for i in range(len(OUTCOMES))}
        if not keep_levels:
            res.close()
        return res
    except Exception:
        res.close()
        raise
