"""The N-API addon executed for real: addon/test/napi_host.c is a minimal Node-API host (this image has no
Node.js) that exports napi_* like the node binary, dlopen()s sift_b200.node, registers it and calls its
functions the way addon/background.js does."""
import os
import subprocess

import numpy as np
import pytest

from sift_b200 import _lib as L, fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADDON = os.path.join(ROOT, "addon")
HOST = os.path.join(ADDON, "test", "napi_host")
NODE = os.path.join(ADDON, "sift_b200.node")


def _run(*args):
    subprocess.check_call(["make", "-C", ADDON], stdout=subprocess.DEVNULL)
    return subprocess.run([HOST, NODE, *map(str, args)], capture_output=True, text=True, timeout=300)


def test_addon_registers_and_reports_its_version():
    r = _run("version")
    assert r.returncode == 0, r.stderr
    assert "exports 18" in r.stdout and "version sift_b200" in r.stdout


def test_addon_create_throws_without_a_gpu():
    """No CPU fallback, and failures surface as thrown JS errors carrying the status name and the message."""
    import torch
    r = _run("create")
    assert r.returncode == 0, r.stderr
    if torch.cuda.is_available():
        assert "created" in r.stdout
    else:
        assert "threw SIFT_ERR_NO_DEVICE" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_addon_detect_equals_the_python_host(engine, tmp_path):
    """Same C ABI underneath: the records the addon hands to JS are byte-identical to the ctypes host's."""
    w, h = 320, 240
    u8 = fixtures.synthetic_u8(w, h, 11)
    raw, out = tmp_path / "img.raw", tmp_path / "kp.bin"
    raw.write_bytes(u8.tobytes())
    r = _run("detect", raw, w, h, 4, 1.6, out)
    assert r.returncode == 0, r.stdout + r.stderr
    want, stats = engine.detect(u8, L.default_params(numberOfOctaves=4, minBlurLevel=1.6))
    got = np.frombuffer(out.read_bytes(), dtype=L.KEYPOINT_DTYPE)
    assert f"count {len(want)} stats.keypoints {len(want)}" in r.stdout
    assert got.tobytes() == want.tobytes() and len(want) > 50
    assert "short buffer threw SIFT_ERR_BAD_ARGS" in r.stdout
    engine.build_scale_space(u8, L.default_params(numberOfOctaves=4, minBlurLevel=1.6))
    rgba, _ = engine.level_preview(L.SIFT_LEVEL_DOG, 0, 1, L.SIFT_PREVIEW_MINMAX)
    checksum = int((rgba.reshape(-1, 4).astype(np.uint64) * np.array([1, 2, 3, 4], dtype=np.uint64)).sum())
    assert f"preview bytes {rgba.size} clamped 1 checksum {checksum}" in r.stdout
    assert f"stages octaves 4 levels 6 dog0_1 {2 * w}x{2 * h} candidates {stats['candidates']} refined {len(want)}" in r.stdout
