"""Engine: one GPU context of the C ABI (include/sift_b200.h), the object the host
mirror modules (sift.py, background.py) forward to.  One Engine per GPU per host
thread, like the reference's single worker (background.js:14-50)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L

_DTYPES = {np.dtype(np.uint8): L.SIFT_U8, np.dtype(np.float32): L.SIFT_F32, np.dtype(np.float64): L.SIFT_F64}


def _as_image(image, rgba: bool = False):
    """Accept a Matrix2D (list of rows) or ndarray; returns (contiguous ndarray, dtype code, w, h)."""
    a = np.asarray(image)
    if rgba:
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 4:
            raise TypeError("RGBA input must be uint8 [h, w, 4]")
        a = np.ascontiguousarray(a)
        return a, L.SIFT_RGBA8, a.shape[1], a.shape[0]
    if a.ndim != 2:
        raise TypeError("image must be 2-D (rows x columns)")
    if a.dtype not in _DTYPES:
        a = a.astype(np.float64)          # JS numbers
    a = np.ascontiguousarray(a)
    return a, _DTYPES[a.dtype], a.shape[1], a.shape[0]


def _f64(m) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(m, dtype=np.float64))
    if a.ndim != 2:
        raise TypeError("Matrix2D must be 2-D")
    return a


def _dp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class Engine:
    def __init__(self, device: int = 0):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.sift_create(device, C.byref(h))
        if rc != L.SIFT_OK:
            raise L.SiftError(rc, (self._lib.sift_last_error(None) or b"").decode())
        self._h = h
        self.device = device

    # -- plumbing
    def close(self):
        if getattr(self, "_h", None):
            self._lib.sift_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int, allow=()):
        if rc != L.SIFT_OK and rc not in allow:
            raise L.SiftError(rc, (self._lib.sift_last_error(self._h) or b"").decode())
        return rc

    @property
    def handle(self):
        return self._h

    @property
    def stream(self) -> int:
        return int(self._lib.sift_stream(self._h) or 0)

    @property
    def kernel_launches(self) -> int:
        return int(self._lib.sift_kernel_launches(self._h))

    @property
    def pyramid_serial(self) -> int:
        """Generation of the pyramid the stage calls read; changes whenever any call rebuilds or invalidates it."""
        return int(self._lib.sift_pyramid_serial(self._h))

    @property
    def transfer_bytes(self):
        """(host->device, device->host) bytes copied by the detect calls so far, counted where the copies are issued."""
        a, b = C.c_uint64(), C.c_uint64()
        self._lib.sift_transfer_bytes(self._h, C.byref(a), C.byref(b))
        return int(a.value), int(b.value)

    def synchronize(self):
        self._check(self._lib.sift_synchronize(self._h))

    def flush(self):
        """Order every device-resident detection issued so far before later work on `stream` (non-blocking)."""
        self._check(self._lib.sift_flush(self._h))

    def set_lanes(self, n: int):
        """Frames in flight for detect_device / detect_batch (1..4)."""
        self._check(self._lib.sift_set_lanes(self._h, int(n)))

    def set_keep_gaussian(self, keep: bool):
        """False: keypoint-only mode, the Gaussian levels are computed but not written to HBM."""
        self._check(self._lib.sift_set_keep_gaussian(self._h, 1 if keep else 0))

    def set_profiling(self, enabled: bool):
        self._check(self._lib.sift_set_profiling(self._h, 1 if enabled else 0))

    def get_profile(self) -> dict:
        """{kind: (ms, launch groups)} recorded since the last call (synchronises)."""
        n = len(L.PROF_KINDS)
        ms = (C.c_float * n)()
        cnt = (C.c_int * n)()
        self._check(self._lib.sift_get_profile(self._h, ms, cnt, n))
        return {L.PROF_KINDS[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}

    # -- fused detect
    def detect(self, image, params: L.Params | None = None, rgba: bool = False, capacity: int | None = None,
               **overrides):
        """Full detection on a host image. Returns (structured ndarray of KEYPOINT_DTYPE, stats dict)."""
        a, dt, w, h = _as_image(image, rgba)
        prm = params if params is not None else L.default_params(**overrides)
        cap = capacity if capacity is not None else max(4096, (w * h) // 64)
        while True:
            out = np.zeros(cap, dtype=L.KEYPOINT_DTYPE)
            n = C.c_int()
            st = L.Stats()
            rc = self._lib.sift_detect(self._h, a.ctypes.data, dt, w, h, 0, C.byref(prm), out.ctypes.data, cap,
                                       C.byref(n), C.byref(st))
            if rc == L.SIFT_ERR_CAPACITY and capacity is None and n.value > cap:
                cap = n.value            # the required count; any other capacity failure raises below
                continue
            self._check(rc)
            return out[:n.value], st.as_dict()

    def detect_raw(self, ptr: int, dtype: int, w: int, h: int, pitch: int, prm: L.Params, out_ptr: int, cap: int):
        """sift_detect on raw host pointers (pinned buffers owned by the caller). Returns (n, Stats)."""
        n = C.c_int()
        st = L.Stats()
        self._check(self._lib.sift_detect(self._h, ptr, dtype, w, h, pitch, C.byref(prm), out_ptr, cap, C.byref(n),
                                          C.byref(st)))
        return n.value, st

    def detect_device(self, d_image: int, dtype: int, w: int, h: int, pitch: int, prm: L.Params, d_out: int,
                      cap: int, d_count: int, ordered: bool = False):
        """Asynchronous device-resident detect on sift_stream(); pointers are device addresses."""
        self._check(self._lib.sift_detect_device(self._h, d_image, dtype, w, h, pitch, C.byref(prm), d_out, cap,
                                                 d_count, 1 if ordered else 0))

    def detect_batch_raw(self, ptr: int, dtype: int, w: int, h: int, pitch: int, stride: int, n_images: int,
                         prm: L.Params, out_ptr: int, cap: int, offsets_ptr: int):
        """sift_detect_batch on raw host pointers (pinned buffers owned by the caller). Returns Stats."""
        st = L.Stats()
        self._check(self._lib.sift_detect_batch(self._h, ptr, dtype, w, h, pitch, stride, n_images, C.byref(prm),
                                                out_ptr, cap, C.cast(offsets_ptr, C.POINTER(C.c_int)), C.byref(st)))
        return st

    def detect_batch(self, images, params: L.Params | None = None, capacity: int | None = None, **overrides):
        """images: ndarray [n, h, w] (u8 / f32 / f64). Returns (keypoints, offsets[n+1], stats)."""
        a = np.ascontiguousarray(images)
        if a.ndim != 3 or a.dtype not in _DTYPES:
            raise TypeError("batch must be [n, h, w] of uint8 / float32 / float64")
        n_img, h, w = a.shape
        prm = params if params is not None else L.default_params(**overrides)
        cap = capacity if capacity is not None else max(4096, n_img * max(4096, (w * h) // 64))
        out = np.zeros(cap, dtype=L.KEYPOINT_DTYPE)
        offs = np.zeros(n_img + 1, dtype=np.int32)
        st = L.Stats()
        self._check(self._lib.sift_detect_batch(self._h, a.ctypes.data, _DTYPES[a.dtype], w, h, 0,
                                                a.strides[0], n_img, C.byref(prm), out.ctypes.data, cap,
                                                offs.ctypes.data_as(C.POINTER(C.c_int)), C.byref(st)))
        return out[:offs[-1]], offs, st.as_dict()

    # -- mosaic strips (see mosaic.py)
    def strip_begin(self, params: L.Params, layout: L.StripLayout, rows, rgba: bool = False):
        a, dt, w, h = _as_image(rows, rgba)
        self._strip_cap = max(getattr(self, "_strip_cap", 0), 1 << 16,
                              (layout.own1[0] - layout.own0[0]) * layout.width[0] // 64)
        self._check(self._lib.sift_strip_begin(self._h, C.byref(params), C.byref(layout), a.ctypes.data, dt, 0))

    def strip_seed(self, octave: int) -> int:
        p = C.c_void_p()
        self._check(self._lib.sift_strip_seed(self._h, octave, C.byref(p)))
        return int(p.value)

    def strip_octave(self, octave: int):
        self._check(self._lib.sift_strip_octave(self._h, octave))

    def strip_finish(self, capacity: int | None = None):
        """Scan + refine of the strip's owned rows.  The output capacity follows the strip's previous result (a
        retry repeats the scan and the refinement), starting from one record per 64 owned octave-0 pixels."""
        cap = capacity or getattr(self, "_strip_cap", 1 << 16)
        while True:
            out = self._pinned_records(cap)
            n = C.c_int()
            st = L.Stats()
            rc = self._lib.sift_strip_finish(self._h, out.ctypes.data, cap, C.byref(n), C.byref(st))
            if rc == L.SIFT_ERR_CAPACITY and n.value > cap:
                cap = n.value + n.value // 8
                continue
            self._check(rc)
            self._strip_cap = max(cap, n.value + n.value // 8)
            return out[:n.value].copy() if self._pin is None else out[:n.value], st.as_dict()

    _pin = None

    def _pinned_records(self, cap: int) -> np.ndarray:
        """A keypoint-record array of `cap` entries in page-locked host memory when torch is importable (a strip of a
        gigapixel mosaic downloads 10^5..10^6 records: pageable memory makes that copy several times slower).  Two
        buffers alternate, so a result stays valid until the SECOND strip_finish after it; callers that keep results
        longer copy them (mosaic.merge_keypoints concatenates, which copies)."""
        try:
            import torch
            if self._pin is None:
                self._pin = [None, None, 0]
            self._pin[2] ^= 1
            i = self._pin[2]
            if self._pin[i] is None or self._pin[i].numel() < cap * L.KEYPOINT_DTYPE.itemsize:
                self._pin[i] = torch.empty(cap * L.KEYPOINT_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
            return self._pin[i].numpy()[:cap * L.KEYPOINT_DTYPE.itemsize].view(L.KEYPOINT_DTYPE)
        except Exception:
            self._pin = None
            return np.empty(cap, dtype=L.KEYPOINT_DTYPE)

    def strip_escaped(self) -> np.ndarray:
        """Refinement walks that left this strip's rows in the last strip_finish / strip_resume (WALK_DTYPE)."""
        n = C.c_int()
        out = np.zeros(1 << 16, dtype=L.WALK_DTYPE)
        self._check(self._lib.sift_strip_escaped(self._h, out.ctypes.data, len(out), C.byref(n)))
        return out[:n.value].copy()

    def strip_resume(self, walks: np.ndarray, capacity: int = 1 << 12):
        """Continue walks (WALK_DTYPE) on this strip's pyramid; returns (keypoints, stats of the walks ended here)."""
        w = np.ascontiguousarray(walks, dtype=L.WALK_DTYPE)
        cap = max(capacity, len(w))
        out = np.zeros(cap, dtype=L.KEYPOINT_DTYPE)
        n = C.c_int()
        st = L.Stats()
        self._check(self._lib.sift_strip_resume(self._h, w.ctypes.data, len(w), out.ctypes.data, cap, C.byref(n), C.byref(st)))
        return out[:n.value], st.as_dict()

    # -- stages
    def build_scale_space(self, image, params: L.Params, rgba: bool = False):
        a, dt, w, h = _as_image(image, rgba)
        self._check(self._lib.sift_build_scale_space(self._h, a.ctypes.data, dt, w, h, 0, C.byref(params)))

    def build_dog(self):
        self._check(self._lib.sift_build_dog(self._h))

    def pyramid_info(self):
        o, l = C.c_int(), C.c_int()
        self._check(self._lib.sift_get_pyramid_info(self._h, C.byref(o), C.byref(l)))
        return o.value, l.value

    def octave_size(self, octave: int):
        w, h = C.c_int(), C.c_int()
        self._check(self._lib.sift_get_octave_size(self._h, octave, C.byref(w), C.byref(h)))
        return w.value, h.value

    def blur_level(self, kind: int, octave: int, level: int) -> float:
        v = C.c_double()
        self._check(self._lib.sift_get_blur_level(self._h, kind, octave, level, C.byref(v)))
        return v.value

    def get_level(self, kind: int, octave: int, level: int) -> np.ndarray:
        w, h = self.octave_size(octave)
        out = np.empty((h, w), np.float32)
        self._check(self._lib.sift_get_level(self._h, kind, octave, level, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def level_preview(self, kind: int, octave: int, level: int, mode: int = L.SIFT_PREVIEW_GRAY,
                      coefficient: float = 1.0):
        """RGBA8 display product of a level (ImageData layout [h, w, 4]) and the (min, max) used:
        GRAY = image-utils.js:171-220, SIGMOID = matrix2d.js:148-156, MINMAX = matrix2d.js:169-192."""
        w, h = self.octave_size(octave)
        out = np.empty((h, w, 4), np.uint8)
        mm = np.zeros(2, np.float64)
        self._check(self._lib.sift_get_level_preview(self._h, kind, octave, level, mode, float(coefficient),
                                                     out.ctypes.data, mm.ctypes.data_as(C.POINTER(C.c_double))))
        return out, (float(mm[0]), float(mm[1]))

    def set_pyramid_shape(self, width0: int, height0: int, params: L.Params):
        self._check(self._lib.sift_set_pyramid_shape(self._h, width0, height0, C.byref(params)))

    def set_level(self, kind: int, octave: int, level: int, data):
        a = np.ascontiguousarray(np.asarray(data, dtype=np.float32))
        self._check(self._lib.sift_set_level(self._h, kind, octave, level, a.ctypes.data_as(C.POINTER(C.c_float))))

    def find_candidates(self, want_low_contrast: bool = False, capacity: int | None = None,
                        params: L.Params | None = None):
        cap = capacity if capacity is not None else 1 << 16
        while True:
            out = np.zeros(cap, dtype=L.CANDIDATE_DTYPE)
            n, nl = C.c_int(), C.c_int()
            low = np.zeros(cap if want_low_contrast else 0, dtype=L.CANDIDATE_DTYPE)
            rc = self._lib.sift_find_candidates(self._h, C.byref(params) if params is not None else None,
                                                out.ctypes.data, cap, C.byref(n),
                                                low.ctypes.data if want_low_contrast else None,
                                                cap if want_low_contrast else 0,
                                                C.byref(nl) if want_low_contrast else None)
            if rc == L.SIFT_ERR_CAPACITY and capacity is None and max(n.value, nl.value) > cap:
                cap = max(n.value, nl.value)
                continue
            self._check(rc)
            return out[:n.value], (low[:nl.value] if want_low_contrast else None)

    def refine(self, candidates: np.ndarray, params: L.Params | None = None):
        c = np.ascontiguousarray(candidates, dtype=L.CANDIDATE_DTYPE)
        out = np.zeros(max(1, len(c)), dtype=L.KEYPOINT_DTYPE)
        n = C.c_int()
        st = L.Stats()
        self._check(self._lib.sift_refine(self._h, C.byref(params) if params is not None else None,
                                          c.ctypes.data if len(c) else None, len(c), out.ctypes.data,
                                          len(out), C.byref(n), C.byref(st)))
        return out[:n.value], st.as_dict()

    # -- step functions (float64 Matrix2D in / out)
    def blur_chunk(self, inp: np.ndarray, out: np.ndarray, sigma: float, x1: int, y1: int, x2: int, y2: int):
        assert inp.dtype == np.float64 and out.dtype == np.float64 and out.flags.c_contiguous
        self._check(self._lib.sift_blur_chunk(self._h, _dp(inp), inp.shape[0], inp.shape[1], _dp(out), float(sigma),
                                              x1, y1, x2, y2))

    def subtract_chunk(self, a: np.ndarray, b: np.ndarray, out: np.ndarray, x1: int, y1: int, x2: int, y2: int):
        self._check(self._lib.sift_subtract_chunk(self._h, _dp(a), _dp(b), a.shape[0], a.shape[1], _dp(out),
                                                  x1, y1, x2, y2))

    def find_extremas(self, d0, d1, d2, spo: int, contrast: float = 0.015, prefactor: float = 0.8):
        d0, d1, d2 = _f64(d0), _f64(d1), _f64(d2)
        rows, cols = d1.shape
        cap = max(1, rows * cols)
        cxy = np.zeros((cap, 2), np.int32); cv = np.zeros(cap, np.float64)
        lxy = np.zeros((cap, 2), np.int32); lv = np.zeros(cap, np.float64)
        nc, nl = C.c_int(), C.c_int()
        i32 = C.POINTER(C.c_int32)
        self._check(self._lib.sift_find_extremas(self._h, _dp(d0), _dp(d1), _dp(d2), rows, cols, spo, contrast,
                                                 prefactor, cxy.ctypes.data_as(i32), _dp(cv), cap, C.byref(nc),
                                                 lxy.ctypes.data_as(i32), _dp(lv), cap, C.byref(nl)))
        return (cxy[:nc.value], cv[:nc.value]), (lxy[:nl.value], lv[:nl.value])

    def gradient_hessian(self, dm, dc, dp, m: int, n: int):
        dm, dc, dp = _f64(dm), _f64(dc), _f64(dp)
        g = np.zeros(3); h = np.zeros(9)
        self._check(self._lib.sift_gradient_hessian(self._h, _dp(dm), _dp(dc), _dp(dp), dc.shape[0], dc.shape[1],
                                                    m, n, _dp(g), _dp(h)))
        return g, h.reshape(3, 3)

    def linear_resize(self, m, rate: float) -> np.ndarray:
        a = _f64(m)
        r, c = C.c_int(), C.c_int()
        self._lib.sift_resize_dims(a.shape[0], a.shape[1], float(rate), C.byref(r), C.byref(c))
        out = np.empty((r.value, c.value), np.float64)
        self._check(self._lib.sift_linear_resize(self._h, _dp(a), a.shape[0], a.shape[1], float(rate), _dp(out)))
        return out


_default = {}


def default_engine(device: int = 0) -> Engine:
    """Process-wide engine per device (the reference has exactly one worker, main.js:46)."""
    if device not in _default:
        _default[device] = Engine(device)
    return _default[device]
