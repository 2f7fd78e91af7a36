"""Seeded synthetic inputs (BASELINE.json "procedural blobs+checkerboard").

The reference defines no test images, so every implementation (CUDA engine,
float64 oracle, the reference JS under jsmini/Node) consumes the bytes made here:
grayscale u8, float image = u8 / 255.0 (the reference's range, image-utils.js:114).

Content: 0.5 + 0.15*checker(period 48 px) + max(8, W*H/4096) Gaussian blobs
(centre uniform, sigma in U[1.5,12], amplitude in U[-0.35,0.35]) + seeded +-2
grey-level noise (so no region is exactly flat: an exactly singular Hessian
crashes the reference), clipped and rounded to u8.  The PRNG is an
integer-only splitmix64 so any language can reproduce the stream.
"""
from __future__ import annotations

import numpy as np

_MASK = (1 << 64) - 1


def _splitmix_stream(seed: int, n: int) -> np.ndarray:
    """n successive splitmix64 outputs as uint64 (vectorised: state_i = seed + (i+1)*gamma)."""
    gamma = np.uint64(0x9E3779B97F4A7C15)
    with np.errstate(over="ignore"):
        z = np.uint64(seed & _MASK) + gamma * np.arange(1, n + 1, dtype=np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _uniform(u64: np.ndarray) -> np.ndarray:
    """53-bit uniform doubles in [0,1)."""
    return (u64 >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synthetic_u8(width: int, height: int, seed: int = 1234, blobs: int | None = None,
                 sigma_lo: float = 1.5, sigma_hi: float = 12.0) -> np.ndarray:
    """Return the (height, width) uint8 synthetic frame for ``seed`` (frame i of a batch: 1234+i).
    ``sigma_lo/hi`` narrow the blob sizes for the tiny golden-vector inputs (tests/golden)."""
    nb = blobs if blobs is not None else max(8, (width * height) // 4096)
    r = _uniform(_splitmix_stream(seed, 4 * nb))
    cx = r[0::4] * width
    cy = r[1::4] * height
    sg = sigma_lo + r[2::4] * (sigma_hi - sigma_lo)
    am = -0.35 + r[3::4] * 0.7
    yy, xx = np.mgrid[0:height, 0:width]
    img = 0.5 + 0.15 * ((((xx // 24) + (yy // 24)) & 1).astype(np.float64) * 2.0 - 1.0)
    for i in range(nb):
        rad = int(np.ceil(4.0 * sg[i]))
        x0, x1 = max(0, int(cx[i]) - rad), min(width, int(cx[i]) + rad + 1)
        y0, y1 = max(0, int(cy[i]) - rad), min(height, int(cy[i]) + rad + 1)
        if x0 >= x1 or y0 >= y1:
            continue
        gx = np.exp(-0.5 * ((np.arange(x0, x1) - cx[i]) / sg[i]) ** 2)
        gy = np.exp(-0.5 * ((np.arange(y0, y1) - cy[i]) / sg[i]) ** 2)
        img[y0:y1, x0:x1] += am[i] * np.outer(gy, gx)
    noise = _splitmix_stream(seed ^ 0x5DEECE66D, width * height)
    noise = ((noise >> np.uint64(40)) % np.uint64(5)).astype(np.int64).reshape(height, width) - 2
    u8 = np.clip(np.rint(img * 255.0) + noise, 0, 255).astype(np.uint8)
    return u8


def to_float(u8: np.ndarray) -> np.ndarray:
    """u8 -> float64 in [0,1]: exactly the reference's v / 255.0 (image-utils.js:114)."""
    return u8.astype(np.float64) / 255.0


def write_pgm(path: str, u8: np.ndarray) -> None:
    with open(path, "wb") as f:
        f.write(b"P5\n%d %d\n255\n" % (u8.shape[1], u8.shape[0]))
        f.write(np.ascontiguousarray(u8).tobytes())


def read_pgm(path: str) -> np.ndarray:
    with open(path, "rb") as f:
        data = f.read()
    parts = data.split(b"\n", 3)
    assert parts[0] == b"P5" and parts[2] == b"255"
    w, h = (int(v) for v in parts[1].split())
    return np.frombuffer(parts[3], dtype=np.uint8, count=w * h).reshape(h, w).copy()
