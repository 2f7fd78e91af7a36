"""The fused kernel converts u8 -> v/255.0 with a multiply and one Newton correction instead of a divide
(blur_fused.cu u8_over_255).  Check with exact rational arithmetic that this is the correctly rounded
quotient -- i.e. the reference's `v / 255.0` (image-utils.js:114) -- for every byte value."""
from fractions import Fraction as F


def _fma(a, b, c):
    return float(F(a) * F(b) + F(c))       # Fraction -> float is correctly rounded


def test_u8_over_255_is_exact_for_all_bytes():
    r = 1.0 / 255.0
    for v in range(256):
        x = float(v)
        q = x * r
        q2 = _fma(_fma(-q, 255.0, x), r, q)
        assert q2 == v / 255.0, v
