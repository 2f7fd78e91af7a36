"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol
include/sift_b200.h declares, struct layouts match the header, and the product never reaches the oracle."""
import ctypes as C
import os
import re

import numpy as np

import sift_b200
from sift_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sift_b200.h")
PKG = os.path.join(ROOT, "sift-scale-space-extrema-detection_b200")


def _declared_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"SIFT_API\s+[\w\s\*]+?\b(sift_\w+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = _declared_symbols()
    assert len(syms) >= 25
    lib = C.CDLL(L.LIB_PATH)                      # no compute call: only dlsym
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"libsift_b200.so lacks {missing}"


def test_only_declared_symbols_are_exported():
    """-fvisibility=hidden: nothing but the sift_* C ABI leaves the library."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    ours = {s for s in exported if s.startswith("sift_")}
    assert ours == set(_declared_symbols())


def test_struct_layouts_match_the_header():
    assert C.sizeof(L.Keypoint) == 80 and L.KEYPOINT_DTYPE.itemsize == 80
    assert C.sizeof(L.Candidate) == 24 and L.CANDIDATE_DTYPE.itemsize == 24
    assert C.sizeof(L.Params) == 72
    assert C.sizeof(L.Stats) == 52
    assert C.sizeof(L.StripLayout) == 4 + 7 * 12 * 4
    for name, _ in L.Keypoint._fields_:
        assert getattr(L.Keypoint, name).offset == L.KEYPOINT_DTYPE.fields[name][1]
    # the reference's record fields (background.js:619-628) are all present
    for f in ("octave", "scaleLevel", "localX", "localY", "absoluteSigma", "absoluteX", "absoluteY",
              "interpolatedValue"):
        assert f in L.KEYPOINT_DTYPE.names


def test_default_params_are_the_reference_defaults():
    """worker.js:33-37, sift.js:285,293, background.js:461,480,558,598 -- host-side call, no GPU needed."""
    lib = L.load()
    p = L.Params()
    lib.sift_default_params(C.byref(p))
    assert (p.numberOfOctaves, p.scalesPerOctave, p.minBlurLevel, p.assumedBlur) == (5, 3, 0.8, 0.5)
    assert (p.contrastThreshold, p.preFilterFactor, p.edgeRatio) == (0.015, 0.8, 10.0)
    assert (p.maxIterations, p.offsetBound, p.minInterpixelDistance) == (5, 0.6, 0.5)
    q = L.default_params(numberOfOctaves=4, minBlurLevel=1.6)
    assert q.numberOfOctaves == 4 and q.minBlurLevel == 1.6 and q.scalesPerOctave == 3


def test_resize_dims_host_helper():
    """matrix2d.js:119,124: loop `i += rate` while i < rows."""
    lib = L.load()
    r, c = C.c_int(), C.c_int()
    assert lib.sift_resize_dims(5, 7, 0.5, C.byref(r), C.byref(c)) == 0 and (r.value, c.value) == (10, 14)
    assert lib.sift_resize_dims(5, 7, 2.0, C.byref(r), C.byref(c)) == 0 and (r.value, c.value) == (3, 4)
    assert lib.sift_resize_dims(5, 7, 0.0, C.byref(r), C.byref(c)) == L.SIFT_ERR_BAD_ARGS


def test_creation_fails_loudly_without_a_device():
    """No CPU fallback: on a box without an sm_100 GPU the engine refuses to exist."""
    import torch
    if torch.cuda.is_available():
        return
    try:
        sift_b200.Engine(0)
    except sift_b200.SiftError as e:
        assert e.status in (L.SIFT_ERR_NO_DEVICE, L.SIFT_ERR_CUDA)
    else:
        raise AssertionError("Engine() succeeded without a GPU")


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under the package (or the addon) may reference it."""
    bad = []
    for base in (PKG, os.path.join(ROOT, "addon")):
        for dp, _, files in os.walk(base):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".c", ".h", ".js", ".cc", ".cpp")):
                    text = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"^\s*(import|from)\s+oracle\b|sift_oracle|libsift_oracle", text, re.M):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_fixture_generator_is_reproducible():
    from sift_b200 import fixtures
    a = fixtures.synthetic_u8(64, 48, 1234)
    b = fixtures.synthetic_u8(64, 48, 1234)
    c = fixtures.synthetic_u8(64, 48, 1235)
    assert a.dtype == np.uint8 and a.shape == (48, 64)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    assert int(a.astype(np.int64).sum()) == int(np.frombuffer(a.tobytes(), np.uint8).astype(np.int64).sum())


def test_napi_addon_builds_and_binds_only_the_c_abi():
    """addon/sift_b200.node: exports the N-API registration symbol, imports only napi_* (resolved from the node
    executable at load time) and sift_* symbols that include/sift_b200.h declares."""
    import subprocess
    addon = os.path.join(ROOT, "addon")
    subprocess.check_call(["make", "-C", addon], stdout=subprocess.DEVNULL)
    out = subprocess.run(["nm", "-D", os.path.join(addon, "sift_b200.node")], capture_output=True, text=True).stdout
    defined = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    undefined = {ln.split()[-1] for ln in out.splitlines() if " U " in ln}
    assert "napi_register_module_v1" in defined
    used = {s for s in undefined if s.startswith("sift_")}
    assert used and used <= set(_declared_symbols())
    assert len({s for s in undefined if s.startswith("napi_")}) >= 15
    foreign = {s.split("@")[0] for s in undefined if not s.startswith(("sift_", "napi_", "_"))}
    assert foreign <= {"memset", "memcpy", "free", "malloc", "strlen", "snprintf"}, foreign


def test_js_drop_in_modules_keep_the_reference_export_names():
    """addon/sift.js and addon/background.js export the names src/sift.js and background.js define."""
    sift_js = open(os.path.join(ROOT, "addon", "sift.js")).read()
    for name in ("SIFT_blurMatrix2DChunk", "SIFT_subtractMatrix2DChunk", "SIFT_findExtremas",
                 "SIFT_generateGradientVector", "SIFT_generateHessianMatrix"):
        assert re.search(rf"export function {name}\(", sift_js), name
    bg_js = open(os.path.join(ROOT, "addon", "background.js")).read()
    for name in ("computeGaussianScaleSpace", "computeDifferenceOfGaussians", "findCandidateKeypoints",
                 "refineCandidateKeypoints", "detect"):
        assert re.search(rf"export function {name}\(", bg_js), name
    # the JS glue parses under the test interpreter's tokenizer (syntax smoke; it cannot run without Node)
    from oracle import jsmini
    for f in ("sift.js", "background.js", "native.js"):
        assert len(jsmini.tokenize(open(os.path.join(ROOT, "addon", f)).read())) > 100
