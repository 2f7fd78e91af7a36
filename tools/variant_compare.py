"""Bit-for-bit comparison of two kernel variants (environment knobs are read once per process, so each side runs in
its own interpreter): every Gaussian / DoG level of every octave and the ordered keypoint records.
usage: python tools/variant_compare.py "KNOB_A[=v][+KNOB..]" "KNOB_B.." [WxHxOCT ...]      ('' = default path)"""
import hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys, json, hashlib
sys.path.insert(0, %(root)r)
import numpy as np
import sift_b200
from sift_b200 import _lib as L, fixtures
eng = sift_b200.Engine(0)
out = {}
for spec in %(specs)r:
    w, h, no = (int(v) for v in spec.split("x"))
    u8 = fixtures.synthetic_u8(w, h, 77)
    prm = L.default_params(numberOfOctaves=no, minBlurLevel=1.6)
    eng.build_scale_space(u8, prm)
    sig = hashlib.sha256()
    for o in range(no):
        for kind, n in ((L.SIFT_LEVEL_GAUSSIAN, 6), (L.SIFT_LEVEL_DOG, 5)):
            for s in range(n):
                sig.update(np.ascontiguousarray(eng.get_level(kind, o, s)).tobytes())
    kps, st = eng.detect(u8, prm)
    out[spec] = [sig.hexdigest(), hashlib.sha256(np.ascontiguousarray(kps).tobytes()).hexdigest(), int(len(kps))]
print("RESULT " + json.dumps(out))
"""


def run(knobs, specs):
    env = dict(os.environ)
    for part in knobs.split("+"):
        if part:
            name, _, val = part.partition("=")
            env[name] = val or "1"
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "specs": specs}], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-3000:])
        sys.exit(1)
    return json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])


if __name__ == "__main__":
    a, b = sys.argv[1], sys.argv[2]
    specs = sys.argv[3:] or ["1920x1080x4", "97x61x3", "5x37x2", "640x333x4", "33x32x2"]
    ra, rb = run(a, specs), run(b, specs)
    bad = 0
    for sp in specs:
        same = ra[sp] == rb[sp]
        bad += not same
        print(f"{sp:14s} levels {'same' if ra[sp][0] == rb[sp][0] else 'DIFFER'}  records {'same' if ra[sp][1] == rb[sp][1] else 'DIFFER'}  n = {ra[sp][2]} / {rb[sp][2]}")
    print("IDENTICAL" if not bad else f"{bad} case(s) differ", repr(a), "vs", repr(b))
    sys.exit(1 if bad else 0)
