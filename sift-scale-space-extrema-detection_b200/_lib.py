"""ctypes binding of the C ABI in include/sift_b200.h (libsift_b200.so).

This is the Python analogue of the N-API addon described in INTEGRATION.md: the
same entry points, the same structs.  There is no CPU fallback: if the shared
library is missing or no sm_100 device is present, construction fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsift_b200.so")

SIFT_OK, SIFT_ERR_BAD_ARGS, SIFT_ERR_CUDA, SIFT_ERR_CAPACITY, SIFT_ERR_UNSUPPORTED, SIFT_ERR_NO_DEVICE, \
    SIFT_ERR_STATE = range(7)
SIFT_U8, SIFT_F32, SIFT_F64, SIFT_RGBA8 = range(4)
SIFT_PREVIEW_GRAY, SIFT_PREVIEW_SIGMOID, SIFT_PREVIEW_MINMAX = 0, 1, 2
SIFT_LEVEL_GAUSSIAN, SIFT_LEVEL_DOG = 0, 1
PROF_KINDS = ("blur_octave0", "blur_octave1", "blur_high_octaves", "scan", "refine")

STATUS_NAMES = {0: "SIFT_OK", 1: "SIFT_ERR_BAD_ARGS", 2: "SIFT_ERR_CUDA", 3: "SIFT_ERR_CAPACITY",
                4: "SIFT_ERR_UNSUPPORTED", 5: "SIFT_ERR_NO_DEVICE", 6: "SIFT_ERR_STATE"}


class SiftError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class Params(C.Structure):
    _fields_ = [("numberOfOctaves", C.c_int32), ("scalesPerOctave", C.c_int32),
                ("minBlurLevel", C.c_double), ("assumedBlur", C.c_double),
                ("contrastThreshold", C.c_double), ("preFilterFactor", C.c_double),
                ("edgeRatio", C.c_double), ("maxIterations", C.c_int32), ("reserved0", C.c_int32),
                ("offsetBound", C.c_double), ("minInterpixelDistance", C.c_double)]


class Candidate(C.Structure):
    _fields_ = [("octave", C.c_int32), ("scaleLevel", C.c_int32), ("x", C.c_int32), ("y", C.c_int32),
                ("value", C.c_float), ("reserved0", C.c_int32)]


class Keypoint(C.Structure):
    _fields_ = [("octave", C.c_int32), ("scaleLevel", C.c_int32), ("localX", C.c_int32), ("localY", C.c_int32),
                ("absoluteSigma", C.c_double), ("absoluteX", C.c_double), ("absoluteY", C.c_double),
                ("interpolatedValue", C.c_double), ("offset", C.c_float * 3), ("dogValue", C.c_float),
                ("candScale", C.c_int32), ("candX", C.c_int32), ("candY", C.c_int32), ("iterations", C.c_int32)]


class StripLayout(C.Structure):
    """sift_strip_layout: rows of the global octave grids a mosaic strip owns / holds (include/sift_b200.h)."""
    _fields_ = [("octaves", C.c_int32), ("width", C.c_int32 * 12), ("height", C.c_int32 * 12),
                ("own0", C.c_int32 * 12), ("own1", C.c_int32 * 12), ("top", C.c_int32 * 12),
                ("bottom", C.c_int32 * 12), ("halo", C.c_int32 * 12)]


class Stats(C.Structure):
    _fields_ = [("candidates", C.c_int32), ("lowContrastExtrema", C.c_int32), ("keypoints", C.c_int32),
                ("rejLowContrast", C.c_int32), ("rejEdge", C.c_int32), ("rejLeftScale", C.c_int32),
                ("rejLeftRows", C.c_int32), ("rejLeftCols", C.c_int32), ("rejNoConvergence", C.c_int32),
                ("rejSingular", C.c_int32), ("msDevice", C.c_float), ("kernelLaunches", C.c_int32),
                ("leftStrip", C.c_int32)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


KEYPOINT_DTYPE = np.dtype([("octave", "<i4"), ("scaleLevel", "<i4"), ("localX", "<i4"), ("localY", "<i4"),
                           ("absoluteSigma", "<f8"), ("absoluteX", "<f8"), ("absoluteY", "<f8"),
                           ("interpolatedValue", "<f8"), ("offset", "<f4", (3,)), ("dogValue", "<f4"),
                           ("candScale", "<i4"), ("candX", "<i4"), ("candY", "<i4"), ("iterations", "<i4")])
CANDIDATE_DTYPE = np.dtype([("octave", "<i4"), ("scaleLevel", "<i4"), ("x", "<i4"), ("y", "<i4"),
                            ("value", "<f4"), ("reserved0", "<i4")])
# sift_walk: a refinement walk handed from one mosaic strip to another (include/sift_b200.h)
WALK_DTYPE = np.dtype([("octave", "<i4"), ("scaleLevel", "<i4"), ("x", "<i4"), ("y", "<i4"), ("iteration", "<i4"),
                       ("value", "<f4"), ("candScale", "<i4"), ("candX", "<i4"), ("candY", "<i4"), ("reserved0", "<i4")])
assert WALK_DTYPE.itemsize == 40
assert KEYPOINT_DTYPE.itemsize == C.sizeof(Keypoint) == 80
assert CANDIDATE_DTYPE.itemsize == C.sizeof(Candidate) == 24

_VP, _IP, _DP, _FP = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.POINTER(C.c_float)

# name -> (restype, argtypes); every symbol include/sift_b200.h declares
PROTOTYPES = {
    "sift_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "sift_destroy": (None, [_VP]),
    "sift_last_error": (C.c_char_p, [_VP]),
    "sift_version": (C.c_char_p, []),
    "sift_default_params": (None, [C.POINTER(Params)]),
    "sift_synchronize": (C.c_int, [_VP]),
    "sift_stream": (_VP, [_VP]),
    "sift_flush": (C.c_int, [_VP]),
    "sift_set_lanes": (C.c_int, [_VP, C.c_int]),
    "sift_set_keep_gaussian": (C.c_int, [_VP, C.c_int]),
    "sift_kernel_launches": (C.c_int64, [_VP]),
    "sift_pyramid_serial": (C.c_uint64, [_VP]),
    "sift_debug_poison": (C.c_int, [_VP, C.c_int]),
    "sift_transfer_bytes": (None, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "sift_set_profiling": (C.c_int, [_VP, C.c_int]),
    "sift_get_profile": (C.c_int, [_VP, _FP, _IP, C.c_int]),
    "sift_detect": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(Params), _VP, C.c_int, _IP,
                              C.POINTER(Stats)]),
    "sift_detect_device": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(Params), _VP,
                                     C.c_int, _VP, C.c_int]),
    "sift_detect_batch": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                    C.POINTER(Params), _VP, C.c_int, _IP, C.POINTER(Stats)]),
    "sift_strip_layout_compute": (C.c_int, [C.POINTER(Params), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.POINTER(StripLayout)]),
    "sift_strip_begin": (C.c_int, [_VP, C.POINTER(Params), C.POINTER(StripLayout), _VP, C.c_int, C.c_size_t]),
    "sift_strip_seed": (C.c_int, [_VP, C.c_int, C.POINTER(_VP)]),
    "sift_strip_octave": (C.c_int, [_VP, C.c_int]),
    "sift_mosaic_exchange": (C.c_int, [C.POINTER(_VP), C.c_int, C.c_int]),
    "sift_strip_finish": (C.c_int, [_VP, _VP, C.c_int, _IP, C.POINTER(Stats)]),
    "sift_strip_escaped": (C.c_int, [_VP, _VP, C.c_int, _IP]),
    "sift_strip_resume": (C.c_int, [_VP, _VP, C.c_int, _VP, C.c_int, _IP, C.POINTER(Stats)]),
    "sift_build_scale_space": (C.c_int, [_VP, _VP, C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(Params)]),
    "sift_build_dog": (C.c_int, [_VP]),
    "sift_find_candidates": (C.c_int, [_VP, C.POINTER(Params), _VP, C.c_int, _IP, _VP, C.c_int, _IP]),
    "sift_refine": (C.c_int, [_VP, C.POINTER(Params), _VP, C.c_int, _VP, C.c_int, _IP, C.POINTER(Stats)]),
    "sift_get_pyramid_info": (C.c_int, [_VP, _IP, _IP]),
    "sift_get_octave_size": (C.c_int, [_VP, C.c_int, _IP, _IP]),
    "sift_get_blur_level": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _DP]),
    "sift_get_level": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _FP]),
    "sift_get_level_preview": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_void_p, _DP]),
    "sift_set_pyramid_shape": (C.c_int, [_VP, C.c_int, C.c_int, C.POINTER(Params)]),
    "sift_set_level": (C.c_int, [_VP, C.c_int, C.c_int, C.c_int, _FP]),
    "sift_blur_chunk": (C.c_int, [_VP, _DP, C.c_int, C.c_int, _DP, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int]),
    "sift_subtract_chunk": (C.c_int, [_VP, _DP, _DP, C.c_int, C.c_int, _DP, C.c_int, C.c_int, C.c_int, C.c_int]),
    "sift_find_extremas": (C.c_int, [_VP, _DP, _DP, _DP, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double,
                                     C.POINTER(C.c_int32), _DP, C.c_int, _IP, C.POINTER(C.c_int32), _DP, C.c_int,
                                     _IP]),
    "sift_gradient_hessian": (C.c_int, [_VP, _DP, _DP, _DP, C.c_int, C.c_int, C.c_int, C.c_int, _DP, _DP]),
    "sift_resize_dims": (C.c_int, [C.c_int, C.c_int, C.c_double, _IP, _IP]),
    "sift_linear_resize": (C.c_int, [_VP, _DP, C.c_int, C.c_int, C.c_double, _DP]),
}

_lib = None


def load():
    """dlopen libsift_b200.so and type every exported symbol.  Fails loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} not found: run __graft_entry__.build() (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def default_params(**overrides) -> Params:
    p = Params()
    load().sift_default_params(C.byref(p))
    for k, v in overrides.items():
        if k not in dict((n, None) for n, _ in Params._fields_):
            raise TypeError(f"unknown parameter {k!r}")
        setattr(p, k, v)
    return p
