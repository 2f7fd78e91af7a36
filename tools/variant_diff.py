"""Numerical comparison of two kernel variants whose sums run in a different order (not bit-identical by design):
largest absolute difference of every Gaussian / DoG level, and the keypoint sets side by side.
usage: python tools/variant_diff.py "KNOB_A[=v][+KNOB..]" "KNOB_B.." [WxHxOCT ...]      ('' = default path)"""
import os, subprocess, sys, tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r"""
import sys
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
    for o in range(no):
        for kind, n, nm in ((L.SIFT_LEVEL_GAUSSIAN, 6, "g"), (L.SIFT_LEVEL_DOG, 5, "d")):
            for s in range(n):
                out["%%s_%%s%%d_%%d" %% (spec, nm, o, s)] = np.ascontiguousarray(eng.get_level(kind, o, s))
    kps, st = eng.detect(u8, prm)
    out[spec + "_kp"] = kps
np.savez(%(path)r, **out)
"""


def run(knobs, specs, path):
    env = dict(os.environ)
    for part in knobs.split("+"):
        if part:
            name, _, val = part.partition("=")
            env[name] = val or "1"
    r = subprocess.run([sys.executable, "-c", CHILD % {"root": ROOT, "specs": specs, "path": path}], env=env, capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stdout[-2000:], r.stderr[-3000:])
        sys.exit(1)
    return np.load(path)


if __name__ == "__main__":
    a, b = sys.argv[1], sys.argv[2]
    specs = sys.argv[3:] or ["1920x1080x4", "97x61x3", "5x37x2", "640x333x4", "33x32x2"]
    with tempfile.TemporaryDirectory() as td:
        ra, rb = run(a, specs, td + "/a.npz"), run(b, specs, td + "/b.npz")
        worst = 0.0
        for sp in specs:
            dmax, where, nbits = 0.0, "", 0
            for k in ra.files:
                if not k.startswith(sp + "_") or k.endswith("_kp"):
                    continue
                x, y = ra[k], rb[k]
                if not (np.isfinite(x).all() and np.isfinite(y).all()):
                    print(sp, k, "NON-FINITE values"); dmax = float("inf"); where = k; break
                d = float(np.abs(x.astype(np.float64) - y.astype(np.float64)).max()) if x.size else 0.0
                nbits += int((x != y).sum())
                if d > dmax: dmax, where = d, k
            ka, kb = ra[sp + "_kp"], rb[sp + "_kp"]
            same = len(ka) == len(kb) and ka.tobytes() == kb.tobytes()
            print(f"{sp:14s} max |diff| {dmax:.3e} at {where:24s} differing values {nbits:8d}  keypoints {len(ka)} / {len(kb)} {'identical' if same else 'differ'}")
            worst = max(worst, dmax)
        print("WORST", worst)
        sys.exit(0 if worst < 2e-7 else 1)
