"""TEST INFRASTRUCTURE ONLY -- a small JavaScript interpreter, enough ECMAScript to run the UNMODIFIED
reference sources (background.js, src/sift.js, src/matrix2d.js, src/image-utils.js, src/worker.js).

Why it exists: the reference is browser JavaScript and this image has no JS engine (no node / deno / d8 /
quickjs), and the reference ships no golden vectors.  To pin the oracle to the reference ITSELF rather
than to a reading of it, `oracle/make_golden.py` runs the reference's own files through this interpreter
on tiny inputs and commits the outputs under tests/golden/.

Scope: strict-mode ES modules with function / arrow functions, let / const / var, array and object
destructuring (defaults, renames), template literals, spread, for / for-of / while / switch / if,
the usual operators, Array / Number / Math built-ins the reference touches.  JS Numbers are Python
floats/ints (identical IEEE-754 double arithmetic while integers stay below 2^53); Math.exp/pow/sqrt
are libm's (V8 may differ in the last ulp, which is far below every tolerance used).
Not a general engine: no prototypes, classes, getters, generators, async, regex, labels or `this`.
"""
from __future__ import annotations

import math
import os
import re

# ------------------------------------------------------------------ values


class _Undefined:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


UNDEF = _Undefined()


class JSObject(dict):
    """Plain object: property bag."""
    __slots__ = ()


class JSError(Exception):
    pass


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


class _Return(Exception):
    def __init__(self, value):
        self.value = value


def js_number_to_string(x) -> str:
    if isinstance(x, bool):
        return "true" if x else "false"
    if isinstance(x, int):
        return str(x)
    if x != x:
        return "NaN"
    if x in (math.inf, -math.inf):
        return "Infinity" if x > 0 else "-Infinity"
    if x == int(x) and abs(x) < 1e21:
        return str(int(x))
    return repr(x)      # shortest round-trip digits, like Number.prototype.toString


def to_str(v) -> str:
    if v is None:
        return "null"
    if v is UNDEF:
        return "undefined"
    if isinstance(v, str):
        return v
    if isinstance(v, (int, float)):
        return js_number_to_string(v)
    if isinstance(v, list):
        return ",".join("" if e is None or e is UNDEF else to_str(e) for e in v)
    if isinstance(v, dict):
        return "[object Object]"
    return str(v)


def to_num(v):
    if isinstance(v, (int, float)):
        return v
    if v is None:
        return 0
    if v is UNDEF:
        return math.nan
    if isinstance(v, str):
        s = v.strip()
        if s == "":
            return 0
        try:
            if s in ("Infinity", "+Infinity"):
                return math.inf
            if s == "-Infinity":
                return -math.inf
            return float(s)
        except ValueError:
            return math.nan
    return math.nan


def truthy(v) -> bool:
    if v is None or v is UNDEF:
        return False
    if isinstance(v, float):
        return v == v and v != 0.0
    if isinstance(v, (bool, int, str)):
        return bool(v)
    return True     # objects, arrays, functions


def strict_eq(a, b) -> bool:
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        if isinstance(a, bool) != isinstance(b, bool):
            return False
        return a == b
    if a is None or b is None or a is UNDEF or b is UNDEF:
        return a is b
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    return a is b


def js_round(x):
    """Math.round: nearest integer, ties towards +Infinity."""
    if x != x or x in (math.inf, -math.inf):
        return x
    f = math.floor(x)
    return f + 1 if x - f >= 0.5 else f


def _div(a, b):
    try:
        return a / b
    except ZeroDivisionError:
        if a != a or a == 0:
            return math.nan
        neg = (a < 0) != (math.copysign(1.0, b) < 0)
        return -math.inf if neg else math.inf


def _add(a, b):
    if isinstance(a, (int, float)) and isinstance(b, (int, float)):
        return a + b
    if isinstance(a, (str, list, dict)) or isinstance(b, (str, list, dict)):
        return to_str(a) + to_str(b)
    return to_num(a) + to_num(b)


def _pow(a, b):
    try:
        r = math.pow(a, b)
    except OverflowError:
        return math.inf
    except ValueError:
        return math.nan
    if isinstance(a, int) and isinstance(b, int) and b >= 0 and abs(r) < 2 ** 53:
        return int(r)
    return r


def _exp(x):
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


class Uint8ClampedArray:
    """ImageData.data: stores round-half-even clamped bytes like the typed array does."""

    def __init__(self, n):
        self.buf = bytearray(int(n) if not isinstance(n, (list, bytearray, bytes)) else n)

    def get(self, i):
        i = int(i)
        return self.buf[i] if 0 <= i < len(self.buf) else UNDEF

    def set(self, i, v):
        i = int(i)
        if 0 <= i < len(self.buf):
            v = to_num(v)
            if v != v:
                v = 0
            v = min(255.0, max(0.0, float(v)))
            self.buf[i] = int(round(v))       # Python round == round-half-even == ToUint8Clamp

    @property
    def length(self):
        return len(self.buf)

    def slice(self, a=0, b=None):
        out = Uint8ClampedArray(0)
        out.buf = self.buf[int(a):(len(self.buf) if b is None or b is UNDEF else int(b))]
        return out


# ------------------------------------------------------------------ tokenizer

_PUNCT = ["===", "!==", "...", "**=", ">>>", "=>", "==", "!=", "<=", ">=", "&&", "||", "??", "++", "--", "+=", "-=",
          "*=", "/=", "%=", "**", "<<", ">>", "{", "}", "(", ")", "[", "]", ";", ",", "<", ">", "+", "-", "*", "/", "%",
          "!", "=", "?", ":", ".", "&", "|", "^", "~"]
_NUM = re.compile(r"0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?")
_ID = re.compile(r"[A-Za-z_$][A-Za-z0-9_$]*")
_KEYWORDS = {"let", "const", "var", "function", "return", "if", "else", "for", "while", "do", "break", "continue",
             "switch", "case", "default", "new", "typeof", "null", "undefined", "true", "false", "import", "export",
             "from", "of", "in", "throw", "try", "catch", "finally", "void", "delete", "instanceof"}


def tokenize(src: str):
    toks = []
    i, n = 0, len(src)
    while i < n:
        c = src[i]
        if c in " \t\r\n":
            i += 1
            continue
        if src.startswith("//", i):
            j = src.find("\n", i)
            i = n if j < 0 else j
            continue
        if src.startswith("/*", i):
            j = src.find("*/", i + 2)
            if j < 0:
                raise JSError("unterminated comment")
            i = j + 2
            continue
        if c in "'\"":
            j = i + 1
            out = []
            while src[j] != c:
                if src[j] == "\\":
                    j += 1
                    out.append({"n": "\n", "t": "\t", "r": "\r", "0": "\0"}.get(src[j], src[j]))
                else:
                    out.append(src[j])
                j += 1
            toks.append(("str", "".join(out), i))
            i = j + 1
            continue
        if c == "`":
            j = i + 1
            parts, cur = [], []
            while src[j] != "`":
                if src[j] == "\\":
                    cur.append({"n": "\n", "t": "\t"}.get(src[j + 1], src[j + 1]))
                    j += 2
                elif src.startswith("${", j):
                    depth, k = 1, j + 2
                    while depth:
                        if src[k] == "{":
                            depth += 1
                        elif src[k] == "}":
                            depth -= 1
                        k += 1
                    parts.append("".join(cur))
                    cur = []
                    parts.append(tokenize(src[j + 2:k - 1]))
                    j = k
                else:
                    cur.append(src[j])
                    j += 1
            parts.append("".join(cur))
            toks.append(("tpl", parts, i))
            i = j + 1
            continue
        m = _NUM.match(src, i)
        if m and (c.isdigit() or (c == "." and i + 1 < n and src[i + 1].isdigit())):
            t = m.group(0)
            if t.lower().startswith("0x"):
                v = int(t, 16)
            elif re.fullmatch(r"\d+", t):
                v = int(t)
            else:
                v = float(t)
            toks.append(("num", v, i))
            i = m.end()
            continue
        m = _ID.match(src, i)
        if m:
            t = m.group(0)
            toks.append(("kw" if t in _KEYWORDS else "id", t, i))
            i = m.end()
            continue
        for p in _PUNCT:
            if src.startswith(p, i):
                toks.append(("p", p, i))
                i += len(p)
                break
        else:
            raise JSError(f"unexpected character {c!r} at {i}")
    toks.append(("eof", None, n))
    return toks


# ------------------------------------------------------------------ parser -> AST tuples


class Parser:
    def __init__(self, toks):
        self.t = toks
        self.i = 0

    # -- helpers
    def peek(self, k=0):
        return self.t[self.i + k]

    def at(self, kind, val=None, k=0):
        t = self.t[self.i + k]
        return t[0] == kind and (val is None or t[1] == val)

    def atp(self, val, k=0):
        return self.at("p", val, k)

    def atk(self, val, k=0):
        return self.at("kw", val, k)

    def eat(self, kind, val=None):
        t = self.t[self.i]
        if t[0] != kind or (val is not None and t[1] != val):
            raise JSError(f"expected {kind} {val!r}, got {t[:2]} at {t[2]}")
        self.i += 1
        return t

    def opt(self, kind, val):
        if self.at(kind, val):
            self.i += 1
            return True
        return False

    def semi(self):
        self.opt("p", ";")

    # -- program
    def program(self):
        body = []
        while not self.at("eof"):
            body.append(self.statement(top=True))
        return ("block", body)

    def statement(self, top=False):
        t = self.peek()
        if t[0] == "str" and top and self.atp(";", 1):      # 'use strict';
            self.i += 2
            return ("empty",)
        if t[0] == "kw":
            k = t[1]
            if k == "import":
                return self.import_decl()
            if k == "export":
                self.i += 1
                st = self.statement()
                return ("export", st)
            if k in ("let", "const", "var"):
                st = self.var_decl()
                self.semi()
                return st
            if k == "function":
                self.i += 1
                name = self.eat("id")[1]
                params, body = self.func_rest()
                return ("funcdecl", name, params, body)
            if k == "return":
                self.i += 1
                e = None
                if not self.atp(";") and not self.atp("}"):
                    e = self.expression()
                self.semi()
                return ("return", e)
            if k == "if":
                self.i += 1
                self.eat("p", "(")
                c = self.expression()
                self.eat("p", ")")
                a = self.statement()
                b = None
                if self.opt("kw", "else"):
                    b = self.statement()
                return ("if", c, a, b)
            if k == "for":
                return self.for_stmt()
            if k == "while":
                self.i += 1
                self.eat("p", "(")
                c = self.expression()
                self.eat("p", ")")
                return ("while", c, self.statement())
            if k == "break":
                self.i += 1
                self.semi()
                return ("break",)
            if k == "continue":
                self.i += 1
                self.semi()
                return ("continue",)
            if k == "switch":
                return self.switch_stmt()
            if k == "throw":
                self.i += 1
                e = self.expression()
                self.semi()
                return ("throw", e)
        if self.atp("{"):
            return self.block()
        if self.atp(";"):
            self.i += 1
            return ("empty",)
        e = self.expression()
        self.semi()
        return ("expr", e)

    def block(self):
        self.eat("p", "{")
        body = []
        while not self.atp("}"):
            body.append(self.statement())
        self.eat("p", "}")
        return ("block", body)

    def import_decl(self):
        self.eat("kw", "import")
        names = []
        self.eat("p", "{")
        while not self.atp("}"):
            n = self.eat("id")[1]
            alias = n
            if self.at("id", "as"):
                self.i += 1
                alias = self.eat("id")[1]
            names.append((n, alias))
            if not self.opt("p", ","):
                break
        self.eat("p", "}")
        self.eat("kw", "from")
        path = self.eat("str")[1]
        self.semi()
        return ("import", names, path)

    def var_decl(self):
        kind = self.peek()[1]
        self.i += 1
        decls = []
        while True:
            target = self.binding_target()
            init = None
            if self.opt("p", "="):
                init = self.assignment()
            decls.append((target, init))
            if not self.opt("p", ","):
                break
        return ("var", kind, decls)

    def binding_target(self):
        """identifier | [a, b = d, ...r] | {a, b: c, d = 1, e: f = 2}"""
        if self.atp("["):
            self.i += 1
            elems = []
            while not self.atp("]"):
                if self.atp(","):
                    self.i += 1
                    elems.append(None)
                    continue
                rest = self.opt("p", "...")
                tgt = self.binding_target()
                dflt = self.assignment() if self.opt("p", "=") else None
                elems.append((tgt, dflt, rest))
                if not self.opt("p", ","):
                    break
            self.eat("p", "]")
            return ("apat", elems)
        if self.atp("{"):
            self.i += 1
            props = []
            while not self.atp("}"):
                key = self.peek()
                if key[0] not in ("id", "kw", "str"):
                    raise JSError(f"bad object pattern key {key}")
                self.i += 1
                tgt = ("name", key[1])
                if self.opt("p", ":"):
                    tgt = self.binding_target()
                dflt = self.assignment() if self.opt("p", "=") else None
                props.append((key[1], tgt, dflt))
                if not self.opt("p", ","):
                    break
            self.eat("p", "}")
            return ("opat", props)
        return ("name", self.eat("id")[1])

    def params(self):
        self.eat("p", "(")
        ps = []
        while not self.atp(")"):
            rest = self.opt("p", "...")
            tgt = self.binding_target()
            dflt = self.assignment() if self.opt("p", "=") else None
            ps.append((tgt, dflt, rest))
            if not self.opt("p", ","):
                break
        self.eat("p", ")")
        return ps

    def func_rest(self):
        ps = self.params()
        body = self.block()
        return ps, body

    def for_stmt(self):
        self.eat("kw", "for")
        self.eat("p", "(")
        init = None
        if self.atk("let") or self.atk("const") or self.atk("var"):
            # for (const x of xs)
            if self.peek(2)[0] == "kw" and self.peek(2)[1] == "of" and self.peek(1)[0] == "id":
                kind = self.peek()[1]
                name = self.peek(1)[1]
                self.i += 3
                it = self.expression()
                self.eat("p", ")")
                return ("forof", kind, ("name", name), it, self.statement())
            init = self.var_decl()
        elif not self.atp(";"):
            init = ("expr", self.expression())
        self.eat("p", ";")
        cond = None if self.atp(";") else self.expression()
        self.eat("p", ";")
        upd = None if self.atp(")") else self.expression()
        self.eat("p", ")")
        return ("for", init, cond, upd, self.statement())

    def switch_stmt(self):
        self.eat("kw", "switch")
        self.eat("p", "(")
        disc = self.expression()
        self.eat("p", ")")
        self.eat("p", "{")
        cases = []
        while not self.atp("}"):
            if self.opt("kw", "default"):
                test = None
            else:
                self.eat("kw", "case")
                test = self.expression()
            self.eat("p", ":")
            body = []
            while not (self.atk("case") or self.atk("default") or self.atp("}")):
                body.append(self.statement())
            cases.append((test, body))
        self.eat("p", "}")
        return ("switch", disc, cases)

    # -- expressions
    def expression(self):
        e = self.assignment()
        while self.atp(","):
            self.i += 1
            e = ("seq", e, self.assignment())
        return e

    def is_arrow(self):
        if self.at("id") and self.atp("=>", 1):
            return True
        if not self.atp("("):
            return False
        depth, k = 0, 0
        while True:
            t = self.peek(k)
            if t[0] == "eof":
                return False
            if t[0] == "p" and t[1] in "([{":
                depth += 1
            elif t[0] == "p" and t[1] in ")]}":
                depth -= 1
                if depth == 0:
                    return self.atp("=>", k + 1)
            k += 1

    def arrow(self):
        if self.at("id"):
            ps = [(("name", self.eat("id")[1]), None, False)]
        else:
            ps = self.params()
        self.eat("p", "=>")
        if self.atp("{"):
            body = self.block()
        else:
            body = ("return", self.assignment())
        return ("func", None, ps, body)

    def assignment(self):
        if self.is_arrow():
            return self.arrow()
        left = self.conditional()
        t = self.peek()
        if t[0] == "p" and t[1] in ("=", "+=", "-=", "*=", "/=", "%="):
            self.i += 1
            right = self.assignment()
            if left[0] not in ("id", "member", "index"):
                raise JSError(f"bad assignment target {left[0]} at {t[2]}")
            return ("assign", t[1], left, right)
        return left

    def conditional(self):
        c = self.binary(0)
        if self.atp("?"):
            self.i += 1
            a = self.assignment()
            self.eat("p", ":")
            b = self.assignment()
            return ("cond", c, a, b)
        return c

    _PREC = [("||", "??"), ("&&",), ("|",), ("^",), ("&",), ("===", "!==", "==", "!="), ("<", ">", "<=", ">=", "instanceof"),
             ("<<", ">>", ">>>"), ("+", "-"), ("*", "/", "%")]

    def binary(self, level):
        if level == len(self._PREC):
            return self.exponent()
        left = self.binary(level + 1)
        while True:
            t = self.peek()
            if t[0] in ("p", "kw") and t[1] in self._PREC[level]:
                self.i += 1
                right = self.binary(level + 1)
                left = ("bin", t[1], left, right)
            else:
                return left

    def exponent(self):
        base = self.unary()
        if self.atp("**"):
            self.i += 1
            return ("bin", "**", base, self.exponent())
        return base

    def unary(self):
        t = self.peek()
        if t[0] == "p" and t[1] in ("!", "-", "+", "~"):
            self.i += 1
            return ("un", t[1], self.unary())
        if t[0] == "p" and t[1] in ("++", "--"):
            self.i += 1
            return ("update", t[1], True, self.unary())
        if t[0] == "kw" and t[1] in ("typeof", "void"):
            self.i += 1
            return ("un", t[1], self.unary())
        return self.postfix()

    def postfix(self):
        e = self.call_member()
        t = self.peek()
        if t[0] == "p" and t[1] in ("++", "--"):
            self.i += 1
            return ("update", t[1], False, e)
        return e

    def arguments(self):
        self.eat("p", "(")
        args = []
        while not self.atp(")"):
            if self.opt("p", "..."):
                args.append(("spread", self.assignment()))
            else:
                args.append(self.assignment())
            if not self.opt("p", ","):
                break
        self.eat("p", ")")
        return args

    def call_member(self):
        if self.atk("new"):
            self.i += 1
            callee = self.primary()
            while self.atp("."):
                self.i += 1
                callee = ("member", callee, self.peek()[1])
                self.i += 1
            args = self.arguments() if self.atp("(") else []
            e = ("new", callee, args)
        else:
            e = self.primary()
        while True:
            if self.atp("."):
                self.i += 1
                name = self.peek()
                if name[0] not in ("id", "kw"):
                    raise JSError(f"bad member name {name}")
                self.i += 1
                e = ("member", e, name[1])
            elif self.atp("["):
                self.i += 1
                idx = self.expression()
                self.eat("p", "]")
                e = ("index", e, idx)
            elif self.atp("("):
                e = ("call", e, self.arguments())
            else:
                return e

    def primary(self):
        t = self.peek()
        k = t[0]
        if k == "num" or k == "str":
            self.i += 1
            return ("lit", t[1])
        if k == "tpl":
            self.i += 1
            parts = []
            for p in t[1]:
                if isinstance(p, str):
                    parts.append(("lit", p))
                else:
                    parts.append(Parser(p).expression())
            return ("tpl", parts)
        if k == "id":
            self.i += 1
            return ("id", t[1])
        if k == "kw":
            if t[1] in ("null", "undefined", "true", "false"):
                self.i += 1
                return ("lit", {"null": None, "undefined": UNDEF, "true": True, "false": False}[t[1]])
            if t[1] == "function":
                self.i += 1
                name = self.eat("id")[1] if self.at("id") else None
                ps, body = self.func_rest()
                return ("func", name, ps, body)
            if t[1] == "import" and self.atp(".", 1) and self.peek(2)[1] == "meta":
                self.i += 3                                   # import.meta (ES modules): an object with `url`
                return ("importmeta",)
        if k == "p":
            if t[1] == "(":
                self.i += 1
                e = self.expression()
                self.eat("p", ")")
                return e
            if t[1] == "[":
                self.i += 1
                elems = []
                while not self.atp("]"):
                    if self.opt("p", "..."):
                        elems.append(("spread", self.assignment()))
                    else:
                        elems.append(self.assignment())
                    if not self.opt("p", ","):
                        break
                self.eat("p", "]")
                return ("array", elems)
            if t[1] == "{":
                self.i += 1
                props = []
                while not self.atp("}"):
                    if self.opt("p", "..."):
                        props.append(("spread", self.assignment()))
                    else:
                        key = self.peek()
                        if key[0] not in ("id", "kw", "str", "num"):
                            raise JSError(f"bad object key {key}")
                        self.i += 1
                        kname = key[1] if isinstance(key[1], str) else js_number_to_string(key[1])
                        if self.opt("p", ":"):
                            props.append((kname, self.assignment()))
                        else:
                            props.append((kname, ("id", kname)))
                    if not self.opt("p", ","):
                        break
                self.eat("p", "}")
                return ("object", props)
        raise JSError(f"unexpected token {t[:2]} at {t[2]}")


# ------------------------------------------------------------------ evaluator (AST -> Python closures)


class Scope:
    __slots__ = ("vars", "parent")

    def __init__(self, parent=None):
        self.vars = {}
        self.parent = parent

    def lookup(self, name):
        s = self
        while s is not None:
            v = s.vars
            if name in v:
                return v
            s = s.parent
        raise JSError(f"ReferenceError: {name} is not defined")


def _declares(stmts):
    for s in stmts:
        if s[0] in ("var", "funcdecl") or (s[0] == "export" and s[1][0] in ("var", "funcdecl")):
            return True
    return False


class Compiler:
    def __init__(self, interp):
        self.interp = interp

    # ---- binding patterns
    def bind(self, target):
        kind = target[0]
        if kind == "name":
            name = target[1]

            def b(scope, value):
                scope.vars[name] = value
            return b
        if kind == "apat":
            elems = []
            for e in target[1]:
                if e is None:
                    elems.append(None)
                else:
                    tgt, dflt, rest = e
                    elems.append((self.bind(tgt), self.expr(dflt) if dflt is not None else None, rest))

            def b(scope, value):
                seq = value if isinstance(value, list) else list(value)
                for i, e in enumerate(elems):
                    if e is None:
                        continue
                    bt, d, rest = e
                    if rest:
                        bt(scope, list(seq[i:]))
                        break
                    v = seq[i] if i < len(seq) else UNDEF
                    if v is UNDEF and d is not None:
                        v = d(scope)
                    bt(scope, v)
            return b
        if kind == "opat":
            props = [(k, self.bind(t), self.expr(d) if d is not None else None) for k, t, d in target[1]]

            def b(scope, value):
                if value is None or value is UNDEF:
                    raise JSError("TypeError: cannot destructure null/undefined")
                for k, bt, d in props:
                    v = get_member(value, k)
                    if v is UNDEF and d is not None:
                        v = d(scope)
                    bt(scope, v)
            return b
        raise JSError(f"bad binding target {kind}")

    # ---- functions
    def func(self, node):
        _, name, params, body = node
        binders = [(self.bind(t), self.expr(d) if d is not None else None, rest) for t, d, rest in params]
        run = self.stmt(body, new_scope=False)
        nparams = len(binders)

        def make(scope):
            def fn(*args):
                s = Scope(scope)
                for i in range(nparams):
                    bt, d, rest = binders[i]
                    if rest:
                        bt(s, list(args[i:]))
                        break
                    v = args[i] if i < len(args) else UNDEF
                    if v is UNDEF and d is not None:
                        v = d(s)
                    bt(s, v)
                try:
                    run(s)
                except _Return as r:
                    return r.value
                return UNDEF
            fn.js_name = name
            return fn
        return make

    # ---- statements
    def stmt(self, node, new_scope=True):
        k = node[0]
        if k == "block":
            stmts = node[1]
            hoisted = [(s[1] if s[0] == "funcdecl" else s[1][1], self.func(("func",) + (s[1:] if s[0] == "funcdecl" else s[1][1:])))
                       for s in stmts if s[0] == "funcdecl" or (s[0] == "export" and s[1][0] == "funcdecl")]
            runs = [self.stmt(s) for s in stmts if s[0] != "empty"]
            need = new_scope and _declares(stmts)

            def run(scope):
                s = Scope(scope) if need else scope
                for name, mk in hoisted:
                    s.vars[name] = mk(s)
                for r in runs:
                    r(s)
            return run
        if k == "empty":
            return lambda scope: None
        if k == "expr":
            e = self.expr(node[1])
            return lambda scope: e(scope)
        if k == "var":
            decls = [(self.bind(t), self.expr(i) if i is not None else None) for t, i in node[2]]

            def run(scope):
                for bt, init in decls:
                    bt(scope, init(scope) if init is not None else UNDEF)
            return run
        if k == "funcdecl":
            return lambda scope: None          # hoisted by the enclosing block
        if k == "export":
            inner = node[1]
            run_inner = self.stmt(inner)
            if inner[0] == "funcdecl":
                names = [inner[1]]
            else:
                names = [t[1] for t, _ in inner[2] if t[0] == "name"]
            interp = self.interp

            def run(scope):
                run_inner(scope)
                for n in names:
                    interp.current_exports[n] = scope.lookup(n)[n]
            return run
        if k == "import":
            names, path = node[1], node[2]
            interp = self.interp

            def run(scope):
                exports = interp.load_module(path)
                for n, alias in names:
                    if n not in exports:
                        raise JSError(f"SyntaxError: module {path} does not export {n}")
                    scope.vars[alias] = exports[n]
            return run
        if k == "return":
            e = self.expr(node[1]) if node[1] is not None else None

            def run(scope):
                raise _Return(e(scope) if e is not None else UNDEF)
            return run
        if k == "if":
            c = self.expr(node[1])
            a = self.stmt(node[2])
            b = self.stmt(node[3]) if node[3] is not None else None

            def run(scope):
                if truthy(c(scope)):
                    a(scope)
                elif b is not None:
                    b(scope)
            return run
        if k == "for":
            init = self.stmt(node[1]) if node[1] is not None else None
            cond = self.expr(node[2]) if node[2] is not None else None
            upd = self.expr(node[3]) if node[3] is not None else None
            body = self.stmt(node[4])
            per_iteration = node[1] is not None and node[1][0] == "var" and node[1][1] != "var" and _contains_func(node[4])

            def run(scope):
                s = Scope(scope)
                if init is not None:
                    init(s)
                while cond is None or truthy(cond(s)):
                    if per_iteration:          # let-bindings are fresh per iteration when closures capture them
                        it = Scope(s)
                        it.vars.update(s.vars)
                        try:
                            body(it)
                        except _Break:
                            break
                        except _Continue:
                            pass
                        s.vars.update({k_: it.vars[k_] for k_ in s.vars})
                    else:
                        try:
                            body(s)
                        except _Break:
                            break
                        except _Continue:
                            pass
                    if upd is not None:
                        upd(s)
            return run
        if k == "forof":
            bt = self.bind(node[2])
            it = self.expr(node[3])
            body = self.stmt(node[4])

            def run(scope):
                for v in list(it(scope)):
                    s = Scope(scope)
                    bt(s, v)
                    try:
                        body(s)
                    except _Break:
                        break
                    except _Continue:
                        continue
            return run
        if k == "while":
            c = self.expr(node[1])
            body = self.stmt(node[2])

            def run(scope):
                while truthy(c(scope)):
                    try:
                        body(scope)
                    except _Break:
                        break
                    except _Continue:
                        continue
            return run
        if k == "break":
            def run(scope):
                raise _Break()
            return run
        if k == "continue":
            def run(scope):
                raise _Continue()
            return run
        if k == "switch":
            disc = self.expr(node[1])
            cases = [(self.expr(t) if t is not None else None, [self.stmt(s) for s in body]) for t, body in node[2]]

            def run(scope):
                v = disc(scope)
                s = Scope(scope)
                start = None
                for i, (t, _) in enumerate(cases):
                    if t is not None and strict_eq(v, t(s)):
                        start = i
                        break
                if start is None:
                    for i, (t, _) in enumerate(cases):
                        if t is None:
                            start = i
                            break
                if start is None:
                    return
                try:
                    for _, body in cases[start:]:
                        for r in body:
                            r(s)
                except _Break:
                    pass
            return run
        if k == "throw":
            e = self.expr(node[1])

            def run(scope):
                raise JSError(f"uncaught: {to_str(e(scope))}")
            return run
        raise JSError(f"unsupported statement {k}")

    # ---- expressions
    def expr(self, node):
        k = node[0]
        if k == "lit":
            v = node[1]
            return lambda scope: v
        if k == "id":
            name = node[1]

            def ev(scope):
                s = scope
                while s is not None:
                    v = s.vars
                    if name in v:
                        return v[name]
                    s = s.parent
                raise JSError(f"ReferenceError: {name} is not defined")
            return ev
        if k == "tpl":
            parts = [self.expr(p) for p in node[1]]
            return lambda scope: "".join(to_str(p(scope)) for p in parts)
        if k == "array":
            elems = [(e[0] == "spread", self.expr(e[1] if e[0] == "spread" else e)) for e in node[1]]
            if not any(sp for sp, _ in elems):
                fs = [f for _, f in elems]
                return lambda scope: [f(scope) for f in fs]

            def ev(scope):
                out = []
                for sp, f in elems:
                    if sp:
                        out.extend(f(scope))
                    else:
                        out.append(f(scope))
                return out
            return ev
        if k == "object":
            props = [(None, self.expr(p[1])) if p[0] == "spread" else (p[0], self.expr(p[1])) for p in node[1]]

            def ev(scope):
                o = JSObject()
                for key, f in props:
                    if key is None:
                        o.update(f(scope))
                    else:
                        o[key] = f(scope)
                return o
            return ev
        if k == "func":
            mk = self.func(node)
            return lambda scope: mk(scope)
        if k == "seq":
            a, b = self.expr(node[1]), self.expr(node[2])

            def ev(scope):
                a(scope)
                return b(scope)
            return ev
        if k == "cond":
            c, a, b = self.expr(node[1]), self.expr(node[2]), self.expr(node[3])
            return lambda scope: a(scope) if truthy(c(scope)) else b(scope)
        if k == "un":
            op, a = node[1], self.expr(node[2])
            if op == "!":
                return lambda scope: not truthy(a(scope))
            if op == "-":
                def ev(scope):
                    v = to_num(a(scope))
                    return -float(v) if v == 0 else -v      # -0 must stay a float zero
                return ev
            if op == "+":
                return lambda scope: to_num(a(scope))
            if op == "typeof":
                def ev(scope):
                    try:
                        v = a(scope)
                    except JSError:
                        return "undefined"
                    if v is UNDEF:
                        return "undefined"
                    if isinstance(v, bool):
                        return "boolean"
                    if isinstance(v, (int, float)):
                        return "number"
                    if isinstance(v, str):
                        return "string"
                    if callable(v):
                        return "function"
                    return "object"
                return ev
            if op == "void":
                def ev(scope):
                    a(scope)
                    return UNDEF
                return ev
            raise JSError(f"unsupported unary {op}")
        if k == "bin":
            return self.binop(node)
        if k == "member":
            obj, name = self.expr(node[1]), node[2]
            return lambda scope: get_member(obj(scope), name)
        if k == "index":
            obj, idx = self.expr(node[1]), self.expr(node[2])

            def ev(scope):
                o = obj(scope)
                i = idx(scope)
                if type(o) is list:
                    if type(i) is int:
                        return o[i] if 0 <= i < len(o) else UNDEF
                    if isinstance(i, float) and i == int(i):
                        i = int(i)
                        return o[i] if 0 <= i < len(o) else UNDEF
                    return get_member(o, to_str(i))
                return get_index(o, i)
            return ev
        if k == "call":
            return self.call(node)
        if k == "importmeta":
            interp = self.interp
            return lambda scope: JSObject(url="file://" + interp._dir_stack[-1] + "/")
        if k == "new":
            callee = self.expr(node[1])
            args = self.args(node[2])

            def ev(scope):
                c = callee(scope)
                if not callable(c):
                    raise JSError("TypeError: not a constructor")
                return c(*args(scope))
            return ev
        if k == "assign":
            return self.assign(node)
        if k == "update":
            op, prefix, target = node[1], node[2], node[3]
            delta = 1 if op == "++" else -1
            getter = self.expr(target)
            setter = self.setter(target)

            def ev(scope):
                old = to_num(getter(scope))
                new = old + delta
                setter(scope, new)
                return new if prefix else old
            return ev
        raise JSError(f"unsupported expression {k}")

    def args(self, arg_nodes):
        items = [(a[0] == "spread", self.expr(a[1] if a[0] == "spread" else a)) for a in arg_nodes]
        if not any(sp for sp, _ in items):
            fs = [f for _, f in items]
            if len(fs) == 1:
                f0 = fs[0]
                return lambda scope: (f0(scope),)
            return lambda scope: [f(scope) for f in fs]

        def ev(scope):
            out = []
            for sp, f in items:
                if sp:
                    out.extend(f(scope))
                else:
                    out.append(f(scope))
            return out
        return ev

    def call(self, node):
        callee, args = node[1], self.args(node[2])
        if callee[0] == "member":
            obj, name = self.expr(callee[1]), callee[2]

            def ev(scope):
                o = obj(scope)
                return call_method(o, name, args(scope))
            return ev
        f = self.expr(callee)

        def ev(scope):
            fn = f(scope)
            if not callable(fn):
                raise JSError(f"TypeError: {callee[1] if callee[0] == 'id' else 'expression'} is not a function")
            return fn(*args(scope))
        return ev

    def setter(self, target):
        k = target[0]
        if k == "id":
            name = target[1]

            def st(scope, v):
                scope.lookup(name)[name] = v
            return st
        if k == "member":
            obj, name = self.expr(target[1]), target[2]
            return lambda scope, v: set_member(obj(scope), name, v)
        if k == "index":
            obj, idx = self.expr(target[1]), self.expr(target[2])
            return lambda scope, v: set_index(obj(scope), idx(scope), v)
        raise JSError("bad assignment target")

    def assign(self, node):
        op, target, right = node[1], node[2], self.expr(node[3])
        setter = self.setter(target)
        if op == "=":
            def ev(scope):
                v = right(scope)
                setter(scope, v)
                return v
            return ev
        getter = self.expr(target)
        fn = {"+=": _add, "-=": lambda a, b: to_num(a) - to_num(b), "*=": lambda a, b: to_num(a) * to_num(b),
              "/=": lambda a, b: _div(to_num(a), to_num(b)), "%=": lambda a, b: math.fmod(to_num(a), to_num(b))}[op]

        def ev(scope):
            v = fn(getter(scope), right(scope))
            setter(scope, v)
            return v
        return ev

    def binop(self, node):
        op, a, b = node[1], self.expr(node[2]), self.expr(node[3])
        if op == "||":
            def ev(scope):
                v = a(scope)
                return v if truthy(v) else b(scope)
            return ev
        if op == "&&":
            def ev(scope):
                v = a(scope)
                return b(scope) if truthy(v) else v
            return ev
        if op == "??":
            def ev(scope):
                v = a(scope)
                return b(scope) if v is None or v is UNDEF else v
            return ev
        if op == "instanceof":
            def ev(scope):
                x, c = a(scope), b(scope)
                if isinstance(c, type):                         # host constructors are Python classes
                    return isinstance(x, c)
                if isinstance(c, _ArrayCtor):
                    return type(x) is list
                return False
            return ev
        if op == "===":
            return lambda scope: strict_eq(a(scope), b(scope))
        if op == "!==":
            return lambda scope: not strict_eq(a(scope), b(scope))
        if op in ("==", "!="):
            def loose(x, y):
                if (x is None or x is UNDEF) and (y is None or y is UNDEF):
                    return True
                if isinstance(x, (int, float, str)) and isinstance(y, (int, float, str)) and type(x) is not type(y):
                    return to_num(x) == to_num(y)
                return strict_eq(x, y)
            if op == "==":
                return lambda scope: loose(a(scope), b(scope))
            return lambda scope: not loose(a(scope), b(scope))
        if op == "+":
            def ev(scope):
                x = a(scope)
                y = b(scope)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x + y
                return _add(x, y)
            return ev
        if op == "-":
            def ev(scope):
                x = a(scope)
                y = b(scope)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x - y
                return to_num(x) - to_num(y)
            return ev
        if op == "*":
            def ev(scope):
                x = a(scope)
                y = b(scope)
                tx, ty = type(x), type(y)
                if (tx is float or tx is int) and (ty is float or ty is int):
                    return x * y
                return to_num(x) * to_num(y)
            return ev
        if op == "/":
            return lambda scope: _div(to_num(a(scope)), to_num(b(scope)))
        if op == "%":
            def ev(scope):
                x, y = to_num(a(scope)), to_num(b(scope))
                if y == 0 or x != x or y != y or x in (math.inf, -math.inf):
                    return math.nan
                r = math.fmod(x, y)
                return int(r) if isinstance(x, int) and isinstance(y, int) else r
            return ev
        if op == "**":
            return lambda scope: _pow(to_num(a(scope)), to_num(b(scope)))
        if op in ("<", ">", "<=", ">="):
            import operator
            f = {"<": operator.lt, ">": operator.gt, "<=": operator.le, ">=": operator.ge}[op]

            def ev(scope):
                x = a(scope)
                y = b(scope)
                if isinstance(x, str) and isinstance(y, str):
                    return f(x, y)
                x, y = to_num(x), to_num(y)
                return f(x, y)          # comparisons with NaN are False in Python too
            return ev
        raise JSError(f"unsupported operator {op}")


def _contains_func(node) -> bool:
    if isinstance(node, tuple):
        if node and node[0] == "func":
            return True
        return any(_contains_func(c) for c in node)
    if isinstance(node, list):
        return any(_contains_func(c) for c in node)
    return False


# ------------------------------------------------------------------ property access / built-in methods

def get_index(o, i):
    if isinstance(o, Uint8ClampedArray):
        return o.get(i)
    if hasattr(o, "js_get"):                                 # host objects with indexed elements (typed arrays)
        return o.js_get(i)
    if isinstance(o, str):
        i = int(i)
        return o[i] if 0 <= i < len(o) else UNDEF
    if isinstance(o, dict):
        return o.get(i if isinstance(i, str) else to_str(i), UNDEF)
    if o is None or o is UNDEF:
        raise JSError(f"TypeError: cannot read properties of {to_str(o)} (reading '{to_str(i)}')")
    return UNDEF


def set_index(o, i, v):
    if type(o) is list:
        if isinstance(i, float) and i == int(i):
            i = int(i)
        if isinstance(i, int) and i >= 0:
            if i >= len(o):
                o.extend([UNDEF] * (i + 1 - len(o)))
            o[i] = v
            return
        raise JSError("unsupported array property write")
    if isinstance(o, Uint8ClampedArray):
        o.set(i, v)
        return
    if hasattr(o, "js_set"):
        o.js_set(i, v)
        return
    if isinstance(o, dict):
        o[i if isinstance(i, str) else to_str(i)] = v
        return
    raise JSError(f"TypeError: cannot set properties of {to_str(o)}")


def get_member(o, name):
    if type(o) is list:
        if name == "length":
            return len(o)
        raise JSError(f"unsupported array property {name}")
    if isinstance(o, dict):
        return o.get(name, UNDEF)
    if o is None or o is UNDEF:
        raise JSError(f"TypeError: cannot read properties of {to_str(o)} (reading '{name}')")
    if isinstance(o, str):
        if name == "length":
            return len(o)
        return UNDEF
    if isinstance(o, (int, float)):
        return UNDEF
    try:
        return getattr(o, name)
    except AttributeError:
        return UNDEF


def set_member(o, name, v):
    if isinstance(o, dict):
        o[name] = v
    elif o is None or o is UNDEF or isinstance(o, (int, float, str, list)):
        raise JSError(f"TypeError: cannot set property {name} of {to_str(o)}")
    else:
        setattr(o, name, v)


def call_method(o, name, args):
    if type(o) is list:
        if name == "push":
            o.extend(args)
            return len(o)
        if name == "every":
            f = args[0]
            for i, e in enumerate(o):
                if not truthy(f(e, i, o)):
                    return False
            return True
        if name == "some":
            f = args[0]
            return any(truthy(f(e, i, o)) for i, e in enumerate(o))
        if name == "forEach":
            f = args[0]
            for i, e in enumerate(list(o)):
                f(e, i, o)
            return UNDEF
        if name == "map":
            f = args[0]
            return [f(e, i, o) for i, e in enumerate(o)]
        if name == "filter":
            f = args[0]
            return [e for i, e in enumerate(o) if truthy(f(e, i, o))]
        if name == "slice":
            a = int(args[0]) if len(args) > 0 and args[0] is not UNDEF else 0
            b = int(args[1]) if len(args) > 1 and args[1] is not UNDEF else len(o)
            return o[a:b]
        if name == "indexOf":
            for i, e in enumerate(o):
                if strict_eq(e, args[0]):
                    return i
            return -1
        if name == "join":
            sep = args[0] if args else ","
            return sep.join("" if e is None or e is UNDEF else to_str(e) for e in o)
        raise JSError(f"unsupported Array method {name}")
    if isinstance(o, (int, float)) and not isinstance(o, bool):
        if name == "toString":
            return js_number_to_string(o)
        if name == "toFixed":
            return f"{o:.{int(args[0]) if args else 0}f}"
        raise JSError(f"unsupported Number method {name}")
    if isinstance(o, str):
        if name == "toString":
            return o
        raise JSError(f"unsupported String method {name}")
    fn = get_member(o, name)
    if not callable(fn):
        raise JSError(f"TypeError: {name} is not a function")
    return fn(*args)


# ------------------------------------------------------------------ host globals


class _ArrayCtor:
    """The global Array: `new Array(n)`, Array.isArray, Array.from(iterable | {length}, mapFn)."""

    def __call__(self, *a):
        if len(a) == 1 and isinstance(a[0], (int, float)):
            return [UNDEF] * int(a[0])
        return list(a)

    @staticmethod
    def isArray(v=UNDEF):
        return type(v) is list

    def __getattr__(self, name):
        if name == "from":
            return self._from
        raise AttributeError(name)

    @staticmethod
    def _from(src, fn=None):
        if type(src) is list:
            items = list(src)
        elif isinstance(src, dict):
            items = [UNDEF] * int(to_num(src.get("length", 0)))
        elif hasattr(src, "js_get"):
            items = [src.js_get(i) for i in range(int(src.length))]
        else:
            items = list(src)
        return [fn(v, i) for i, v in enumerate(items)] if fn is not None and fn is not UNDEF else items


class _Number:
    EPSILON = 2.220446049250313e-16
    MAX_SAFE_INTEGER = 9007199254740991
    MIN_SAFE_INTEGER = -9007199254740991
    MAX_VALUE = 1.7976931348623157e308
    MIN_VALUE = 5e-324
    POSITIVE_INFINITY = math.inf
    NEGATIVE_INFINITY = -math.inf
    NaN = math.nan

    def __call__(self, v=0):
        return to_num(v)

    @staticmethod
    def isFinite(v):
        return isinstance(v, (int, float)) and v == v and v not in (math.inf, -math.inf)

    @staticmethod
    def isInteger(v):
        return isinstance(v, (int, float)) and v == v and v not in (math.inf, -math.inf) and v == int(v)


def _math():
    m = JSObject()
    m.update({
        "PI": math.pi, "E": math.e, "SQRT2": math.sqrt(2.0),
        "exp": lambda x: _exp(to_num(x)),
        "pow": lambda a, b: _pow(to_num(a), to_num(b)),
        "sqrt": lambda x: math.sqrt(x) if to_num(x) >= 0 else math.nan,
        "abs": lambda x: abs(to_num(x)),
        "floor": lambda x: x if isinstance(x, int) else (math.floor(x) if x == x and abs(x) != math.inf else x),
        "ceil": lambda x: x if isinstance(x, int) else (math.ceil(x) if x == x and abs(x) != math.inf else x),
        "round": lambda x: x if isinstance(x, int) else js_round(x),
        "max": lambda *a: max(a) if a else -math.inf,
        "min": lambda *a: min(a) if a else math.inf,
        "log": lambda x: math.log(x) if x > 0 else (-math.inf if x == 0 else math.nan),
        "log2": lambda x: math.log2(x) if x > 0 else (-math.inf if x == 0 else math.nan),
        "trunc": lambda x: int(x),
        "sign": lambda x: (x > 0) - (x < 0),
    })
    return m


class Interpreter:
    """Loads ES modules from disk and runs them.  `globals_` are visible to every module (the worker's
    global scope: onmessage, postMessage, console, OffscreenCanvas ...)."""

    def __init__(self, root: str, globals_: dict | None = None):
        self.root = os.path.abspath(root)
        self.global_scope = Scope()
        g = self.global_scope.vars
        g.update({"Math": _math(), "Number": _Number(), "Infinity": math.inf, "NaN": math.nan,
                  "Uint8ClampedArray": Uint8ClampedArray,
                  "Array": _ArrayCtor(),
                  "parseFloat": to_num, "isNaN": lambda v: to_num(v) != to_num(v)})
        if globals_:
            g.update(globals_)
        self.modules = {}
        self.virtual_modules = {}        # specifier -> exports, for host-provided modules ('node:module', ...)
        self.current_exports = None
        self._dir_stack = [self.root]
        self.compiler = Compiler(self)

    def load_module(self, path: str) -> dict:
        if path in self.virtual_modules:
            return self.virtual_modules[path]
        full = os.path.normpath(os.path.join(self._dir_stack[-1], path))
        if full in self.modules:
            return self.modules[full]
        with open(full, "r", encoding="utf-8") as f:
            src = f.read()
        ast = Parser(tokenize(src)).program()
        run = self.compiler.stmt(ast, new_scope=False)
        exports = {}
        self.modules[full] = exports
        saved = self.current_exports
        self.current_exports = exports
        self._dir_stack.append(os.path.dirname(full))
        try:
            run(Scope(self.global_scope))
        finally:
            self._dir_stack.pop()
            self.current_exports = saved
        return exports

    def eval(self, src: str):
        """Evaluate an expression in the global scope (tests)."""
        return self.compiler.expr(Parser(tokenize(src)).expression())(self.global_scope)

    def run(self, src: str):
        run = self.compiler.stmt(Parser(tokenize(src)).program(), new_scope=False)
        self.current_exports = {}
        run(self.global_scope)

    def get_global(self, name):
        return self.global_scope.vars.get(name, UNDEF)


def to_python(v):
    """JS value -> plain Python (lists / dicts / floats) for fixture dumps."""
    if isinstance(v, list):
        return [to_python(e) for e in v]
    if isinstance(v, dict):
        return {k: to_python(e) for k, e in v.items()}
    if v is UNDEF:
        return None
    return v
