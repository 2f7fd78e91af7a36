"""examples/detect_pgm.c: the C ABI used from plain C (no Python, no torch in the process)."""
import os
import subprocess

import numpy as np
import pytest

from sift_b200 import _lib as L, fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "examples", "detect_pgm")


def _build():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "examples")], stdout=subprocess.DEVNULL)


def test_c_example_builds_and_fails_loudly_without_a_gpu(tmp_path):
    import torch
    _build()
    pgm = tmp_path / "a.pgm"
    fixtures.write_pgm(str(pgm), fixtures.synthetic_u8(64, 48, 3))
    r = subprocess.run([EXE, str(pgm), "2", "1.6"], capture_output=True, text=True, timeout=120)
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr and r.stdout == ""
    assert subprocess.run([EXE], capture_output=True).returncode == 2          # usage


@pytest.mark.gpu
def test_c_example_prints_the_python_hosts_records(engine, tmp_path):
    _build()
    u8 = fixtures.synthetic_u8(300, 200, 12)
    pgm = tmp_path / "b.pgm"
    fixtures.write_pgm(str(pgm), u8)
    r = subprocess.run([EXE, str(pgm), "3", "1.6"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    want, _ = engine.detect(u8, L.default_params(numberOfOctaves=3, minBlurLevel=1.6))
    rows = [l.split() for l in r.stdout.strip().splitlines()]
    assert len(rows) == len(want) > 50
    got_i = np.array([[int(v) for v in row[:4]] for row in rows])
    got_f = np.array([[float(v) for v in row[4:]] for row in rows])
    assert np.array_equal(got_i, np.stack([want["octave"], want["scaleLevel"], want["localX"], want["localY"]], axis=1))
    assert np.array_equal(got_f, np.stack([want["absoluteSigma"], want["absoluteX"], want["absoluteY"], want["interpolatedValue"]], axis=1))
