"""Test-only host objects for running addon/*.js under oracle/jsmini.py: the typed arrays, ArrayBuffer and DataView
of ECMAScript (backed by numpy), queueMicrotask, and a `node:module` whose createRequire hands out a Python object
in place of sift_b200.node.  Nothing here is product code."""
from __future__ import annotations

import struct

import numpy as np

from oracle import jsmini

UNDEF = jsmini.UNDEF


class ArrayBuffer:
    def __init__(self, n=0):
        self.data = n if isinstance(n, np.ndarray) else np.zeros(int(n), dtype=np.uint8)

    @property
    def byteLength(self):
        return int(self.data.size)


class TypedArray:
    dtype = np.float64
    clamped = False

    def __init__(self, arg=0, byte_offset=0, length=UNDEF):
        item = np.dtype(self.dtype).itemsize
        if isinstance(arg, ArrayBuffer):
            off = int(byte_offset or 0)
            n = (arg.byteLength - off) // item if length is UNDEF or length is None else int(length)
            self.buffer, self.byteOffset = arg, off
            self.a = arg.data[off:off + n * item].view(self.dtype)
        elif isinstance(arg, (int, float)):
            self.buffer, self.byteOffset = ArrayBuffer(int(arg) * item), 0
            self.a = self.buffer.data.view(self.dtype)
        else:
            vals = arg.a if isinstance(arg, TypedArray) else [jsmini.to_num(v) for v in arg]
            self.buffer, self.byteOffset = ArrayBuffer(len(vals) * item), 0
            self.a = self.buffer.data.view(self.dtype)
            for i, v in enumerate(vals):
                self.js_set(i, v)

    @classmethod
    def of_numpy(cls, arr):
        arr = np.ascontiguousarray(arr, dtype=cls.dtype)
        out = cls(0)
        out.buffer = ArrayBuffer(arr.view(np.uint8).reshape(-1).copy())
        out.a = out.buffer.data.view(cls.dtype)
        return out

    @property
    def length(self):
        return int(self.a.size)

    @property
    def byteLength(self):
        return int(self.a.nbytes)

    def js_get(self, i):
        i = int(i) if isinstance(i, (int, float)) and i == int(i) else -1
        if 0 <= i < self.a.size:
            v = self.a[i]
            return float(v) if np.issubdtype(self.a.dtype, np.floating) else int(v)
        return UNDEF

    def js_set(self, i, v):
        i = int(i)
        if not 0 <= i < self.a.size:
            return
        v = jsmini.to_num(v)
        if self.clamped:
            v = 0.0 if v != v else min(255.0, max(0.0, float(v)))
            self.a[i] = int(round(v))                    # round-half-even == ToUint8Clamp
        elif np.issubdtype(self.a.dtype, np.floating):
            self.a[i] = v
        else:
            self.a[i] = int(v) if v == v else 0

    def set(self, src, offset=0):
        vals = src.a if isinstance(src, TypedArray) else np.array([jsmini.to_num(v) for v in src], dtype=np.float64)
        o = int(offset)
        if o + len(vals) > self.a.size:
            raise jsmini.JSError("RangeError: offset is out of bounds")
        self.a[o:o + len(vals)] = vals

    def subarray(self, a=0, b=UNDEF):
        out = type(self)(0)
        hi = self.a.size if b is UNDEF or b is None else int(b)
        out.a = self.a[int(a):hi]                        # a view, like the real subarray
        out.buffer, out.byteOffset = self.buffer, self.byteOffset + int(a) * self.a.itemsize
        return out

    def slice(self, a=0, b=UNDEF):
        s = self.subarray(a, b)
        return type(self).of_numpy(s.a.copy())

    def __iter__(self):
        return (self.js_get(i) for i in range(self.a.size))


def _typed(name, dtype, clamped=False):
    cls = type(name, (TypedArray,), {"dtype": dtype, "clamped": clamped})
    setattr(cls, "from", classmethod(lambda c, src: c(list(src) if not isinstance(src, list) else src)))
    return cls


Float64Array = _typed("Float64Array", np.float64)
Float32Array = _typed("Float32Array", np.float32)
Int32Array = _typed("Int32Array", np.int32)
Uint8Array = _typed("Uint8Array", np.uint8)
Uint8ClampedArray = _typed("Uint8ClampedArray", np.uint8, clamped=True)


class DataView:
    _FMT = {"Int32": "i", "Uint32": "I", "Float32": "f", "Float64": "d", "Uint8": "B"}

    def __init__(self, buffer, byte_offset=0):
        self.buffer = buffer
        self.off = int(byte_offset or 0)

    def __getattr__(self, name):
        kind = name[3:]
        if name[:3] in ("get", "set") and kind in self._FMT:
            fmt = self._FMT[kind]
            mem = self.buffer.data

            def get(o, little=False):
                return struct.unpack_from(("<" if jsmini.truthy(little) else ">") + fmt, mem, self.off + int(o))[0]

            def put(o, v, little=False):
                v = jsmini.to_num(v)
                struct.pack_into(("<" if jsmini.truthy(little) else ">") + fmt, mem, self.off + int(o),
                                 v if fmt in "fd" else int(v))
            return get if name[:3] == "get" else put
        raise AttributeError(name)


class JSTypeError(Exception):
    def __init__(self, message=""):
        super().__init__(message)
        self.message = message


def make_interpreter(root, native):
    """An interpreter rooted at addon/ whose `require('./sift_b200.node')` returns `native`."""
    microtasks = []
    interp = jsmini.Interpreter(root, {
        "Float64Array": Float64Array, "Float32Array": Float32Array, "Int32Array": Int32Array, "Uint8Array": Uint8Array,
        "Uint8ClampedArray": Uint8ClampedArray, "ArrayBuffer": ArrayBuffer, "DataView": DataView,
        "TypeError": JSTypeError, "Error": JSTypeError, "queueMicrotask": microtasks.append,
    })
    interp.virtual_modules["node:module"] = {"createRequire": lambda url: (lambda spec: native)}

    def drain():
        while microtasks:
            microtasks.pop(0)()
    return interp, drain
