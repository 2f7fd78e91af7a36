"""The JavaScript interpreter that runs the unmodified reference (oracle/jsmini.py) -- language semantics the
reference relies on.  When /root/reference is present (build container only) the reference's own files are
parsed and spot-run as well."""
import math
import os

import pytest

from oracle import jsmini

REF = "/root/reference"


@pytest.fixture()
def js(tmp_path):
    return jsmini.Interpreter(str(tmp_path))


def test_arithmetic_is_ieee_double(js):
    assert js.eval("0.1 + 0.2") == 0.1 + 0.2
    assert js.eval("1 / 3") == 1 / 3
    assert js.eval("7 / 2") == 3.5
    assert js.eval("1 / 0") == math.inf and js.eval("-1 / 0") == -math.inf
    assert math.isnan(js.eval("0 / 0"))
    assert js.eval("2 ** 10") == 1024 and js.eval("7 % 3") == 1
    assert js.eval("Number.EPSILON") == 2.220446049250313e-16
    assert js.eval("Number.MAX_SAFE_INTEGER") == 2 ** 53 - 1


def test_math_round_is_half_up(js):
    assert js.eval("Math.round(2.5)") == 3 and js.eval("Math.round(-2.5)") == -2
    assert js.eval("Math.round(-0.5)") == 0 and js.eval("Math.round(7.4704)") == 7
    assert js.eval("Math.floor(-1.5)") == -2 and js.eval("Math.abs(-3.25)") == 3.25
    assert js.eval("Math.pow(2, 1 / 3)") == math.pow(2, 1 / 3)
    assert js.eval("Math.exp(-0.5)") == math.exp(-0.5) and js.eval("Math.sqrt(2)") == math.sqrt(2)


def test_number_string_round_trip(js):
    """matrix2d.js:129 does Number(x.toString()): value preserving."""
    for v in (0.1, 1 / 3, 5e-324, 1.7976931348623157e308, 123456789.125, 2.0, -0.75):
        js.global_scope.vars["v"] = v
        assert js.eval("Number(v.toString())") == v
    assert js.eval("(2).toString()") == "2" and js.eval("(0.5).toString()") == "0.5"


def test_strict_equality_and_truthiness(js):
    assert js.eval("1 === 1.0") is True and js.eval("null === undefined") is False
    assert js.eval("null == undefined") is True and js.eval("'a' !== 'b'") is True
    assert js.eval("[] === []") is False
    assert js.eval("0 || 5") == 5 and js.eval("3 && 4") == 4 and js.eval("!0") is True
    assert js.eval("NaN > 12.1") is False and js.eval("NaN < 12.1") is False      # background.js:599 (Q6)


def test_closures_destructuring_defaults_spread_templates(js):
    assert js.eval("(() => { let s = 0; const f = v => { s += v; return v; }; f(2); f(3); return s; })()") == 5
    assert js.eval("((a, {b = 2, c: d = 3}) => a + b + d)(1, {c: 5})") == 8
    assert js.eval("(function () { const [x, y] = [1, 2]; const {p, q: r} = {p: 3, q: 4}; return x + y + p + r; })()") == 10
    assert js.eval("((a, b = 7) => a + b)(1)") == 8
    assert js.eval("[...[1, 2], 3].length") == 3
    assert js.eval("`v=${1 + 1} ${'x'}`") == "v=2 x"
    assert js.eval("[1, 2, 3].every(e => e > 0)") is True and js.eval("[1, -2, 3].every(e => e > 0)") is False
    assert js.eval("[4, 5, 6].slice(1, 3)") == [5, 6]


def test_statements(js):
    src = """
    function classify(t) {
      switch (t) {
        case 'a': return 1;
        case 'b': { const z = 2; return z; }
        default: return -1;
      }
    }
    let total = 0;
    for (let i = 0; i < 10; i += 0.5) { if (i === 2) continue; if (i >= 4) break; total += i; }
    const out = [];
    for (const v of [1, 2, 3]) out.push(v * 2);
    let w = 0; while (w < 3) w++;
    """
    js.run(src)
    g = js.global_scope.vars
    assert g["classify"]("a") == 1 and g["classify"]("b") == 2 and g["classify"]("zz") == -1
    assert g["total"] == 0 + 0.5 + 1 + 1.5 + 2.5 + 3 + 3.5
    assert g["out"] == [2, 4, 6] and g["w"] == 3


def test_let_is_per_iteration_for_closures(js):
    js.run("const fs = []; for (let i = 0; i < 3; i++) { fs.push(() => i); }")
    assert [f() for f in js.global_scope.vars["fs"]] == [0, 1, 2]


def test_uint8_clamped_array():
    a = jsmini.Uint8ClampedArray(4)
    a.set(0, 300.0); a.set(1, -5); a.set(2, 127.5); a.set(3, 128.5)
    assert list(a.buf) == [255, 0, 128, 128]            # ToUint8Clamp: round half to even
    assert list(a.slice(1, 3).buf) == [0, 128]


def test_modules_and_exports(tmp_path):
    (tmp_path / "lib").mkdir()
    (tmp_path / "lib" / "m.js").write_text("'use strict';\nexport const K = { A: 'a' };\nexport function twice(x) { return 2 * x; }\n")
    (tmp_path / "main.js").write_text("import { K, twice } from './lib/m.js';\nresult = twice(21) + K.A;\n")
    js = jsmini.Interpreter(str(tmp_path), {"result": None})
    js.load_module("main.js")
    assert js.get_global("result") == "42a"


def test_array_constructor_instanceof_and_host_modules(tmp_path):
    """What addon/*.js needs beyond the reference's subset (tests/test_addon_js_glue.py runs those files)."""
    class Box:                                            # a host class: `new Box(v)`, `x instanceof Box`, indexed reads
        def __init__(self, v=0):
            self.v = v

        def js_get(self, i):
            return self.v + i

        def js_set(self, i, v):
            self.v = v - i

    js = jsmini.Interpreter(str(tmp_path), {"Box": Box})
    assert js.eval("new Array(3).length") == 3 and js.eval("Array.isArray(new Array(2))") is True
    assert js.eval("Array.from([1, 2, 3], (v, i) => v * 10 + i)") == [10, 21, 32]
    assert js.eval("Array.from({ length: 3 }, (_, i) => i * i)") == [0, 1, 4]
    assert js.eval("new Box(5) instanceof Box") is True and js.eval("[] instanceof Box") is False
    assert js.eval("[1] instanceof Array") is True and js.eval("({}) instanceof Array") is False
    assert js.eval("(() => { const b = new Box(5); b[2] = 10; return b[1]; })()") == 9
    assert js.eval("null ?? 7") == 7 and js.eval("0 ?? 7") == 0
    assert js.eval("({ a: 1, ...{ b: 2, a: 3 } }).a") == 3
    js.virtual_modules["host:thing"] = {"answer": lambda: 42}
    (tmp_path / "m.js").write_text("import { answer } from 'host:thing';\nexport const where = import.meta.url;\nexport const v = answer();\n")
    m = js.load_module("m.js")
    assert m["v"] == 42 and m["where"].startswith("file://") and str(tmp_path) in m["where"]


def test_undeclared_assignment_throws_in_strict_modules(tmp_path):
    (tmp_path / "bad.js").write_text("undeclared_global = 1;\n")
    js = jsmini.Interpreter(str(tmp_path))
    with pytest.raises(jsmini.JSError):
        js.load_module("bad.js")


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree only exists in the build container")
def test_reference_sources_load_and_run():
    from oracle import make_golden
    w = make_golden.ReferenceWorker(REF)
    assert callable(w.interp.get_global("onmessage"))                                  # background.js:14
    assert w.types["COMPUTE_GAUSSIAN_SCALE_SPACE"] == "compute-gaussian-scale-space"   # worker.js:6
    inv = w.matrix2d["Matrix2D_get3x3Inverse"]([[2, 0, 0], [0, 4, 0], [0, 0, 8]])
    assert inv == [[0.5, 0, 0], [0, 0.25, 0], [0, 0, 0.125]]
    assert w.matrix2d["Matrix2D_get3x3Inverse"]([[1, 2, 3], [2, 4, 6], [1, 1, 1]]) is None
    up = w.matrix2d["Matrix2D_linearResize"]([[1, 2], [3, 4]], 0.5)
    assert up == [[1, 1, 2, 2], [1, 1, 2, 2], [3, 3, 4, 4], [3, 3, 4, 4]]
