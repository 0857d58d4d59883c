"""jsmini — a small JavaScript interpreter, written to EXECUTE THE REFERENCE'S OWN SOURCE.

TEST INFRASTRUCTURE ONLY (like everything under oracle/).  No JavaScript engine exists in
this image (no node / deno / quickjs, no network), so the reference cannot be run the normal
way.  This module implements the subset of ECMAScript that the three files on the hot path
use, faithfully enough to run them UNMODIFIED:

    /root/reference/src/ola-processor.js
    /root/reference/src/phase-vocoder.js
    fft.js 4.0.3  (module 1 of the browserify bundle /root/reference/www/phase-vocoder.js:2-509)

It is used by tests/golden/generate_golden.py, which runs the reference classes in this
container and commits their outputs as golden vectors; tests/test_oracle_golden.py then
pins oracle/phaze_oracle.c against them.  Nothing here is used on any product path.

Supported: "use strict", var/let/const, function declarations and expressions, classes
(extends, super(), methods, static getters), new, this, prototypes, object and array
literals, if / for / while / break / continue / return / throw, the arithmetic, bitwise,
comparison, logical, conditional, assignment and update operators (including ** and >>>),
Number semantics as IEEE doubles with ToInt32 / ToUint32 for the bitwise family, Array
(including writes to negative "indices", which create named properties), Float32Array /
Int32Array (fill, set, subarray, copyWithin), Math, console.assert.

Scoping: every declaration is function-scoped.  A let/const that would shadow an outer
binding of the same function is rejected at compile time, so this is indistinguishable
from block scoping for the programs accepted.
"""
from __future__ import annotations

import math
import re

import numpy as np

# --------------------------------------------------------------------------------------
# values
# --------------------------------------------------------------------------------------


class _Undefined:
    __slots__ = ()

    def __repr__(self):
        return "undefined"

    def __bool__(self):
        return False


undefined = _Undefined()
null = None


class JSError(Exception):
    def __init__(self, value):
        super().__init__(str(getattr(value, "props", {}).get("message", value)))
        self.value = value


class JSObject:
    __slots__ = ("props", "proto", "getters")

    def __init__(self, proto=None):
        self.props = {}
        self.proto = proto
        self.getters = None

    def get(self, key):
        o = self
        while o is not None:
            if key in o.props:
                return o.props[key]
            if o.getters and key in o.getters:
                return o.getters[key].call(self, [])
            o = o.proto
        return undefined

    def set(self, key, value):
        self.props[key] = value


class JSArray(JSObject):
    """Array: dense list + named properties (a negative 'index' is a named property)."""
    __slots__ = ("items",)

    def __init__(self, items):
        super().__init__(None)
        self.items = items


class JSTyped(JSObject):
    """Float32Array / Int32Array over a numpy array (subarray() shares storage)."""
    __slots__ = ("a",)

    def __init__(self, a):
        super().__init__(None)
        self.a = a


class JSFunction(JSObject):
    __slots__ = ("name", "params", "body", "closure", "is_class", "parent", "native", "hoisted",
                 "ctor_native")

    def __init__(self, name, params, body, closure, native=None):
        super().__init__(None)
        self.name, self.params, self.body, self.closure = name, params, body, closure
        self.is_class, self.parent, self.native = False, None, native
        self.hoisted = ()
        self.ctor_native = None
        self.props["prototype"] = JSObject(None)
        self.props["prototype"].props["constructor"] = self

    def call(self, this, args):
        if self.native is not None:
            return self.native(this, args)
        scope = Scope(self.closure)
        scope.vars["this"] = this
        scope.fn = self
        vs = scope.vars
        for name in self.hoisted:
            vs[name] = undefined
        for i, p in enumerate(self.params):
            vs[p] = args[i] if i < len(args) else undefined
        try:
            self.body(scope)
        except _Return as r:
            return r.value
        return undefined

    def construct(self, args):
        if self.ctor_native is not None:
            return self.ctor_native(args)
        obj = JSObject(self.props["prototype"])
        res = self.call(obj, args)
        return res if isinstance(res, JSObject) else obj


class Scope:
    __slots__ = ("vars", "parent", "fn")

    def __init__(self, parent):
        self.vars = {}
        self.parent = parent
        self.fn = None

    def lookup(self, name):
        s = self
        while s is not None:
            if name in s.vars:
                return s
            s = s.parent
        return None


class _Return(Exception):
    def __init__(self, value):
        self.value = value


class _Break(Exception):
    pass


class _Continue(Exception):
    pass


# --------------------------------------------------------------------------------------
# Number helpers
# --------------------------------------------------------------------------------------

def to_number(v):
    if isinstance(v, float):
        return v
    if isinstance(v, bool):
        return 1.0 if v else 0.0
    if isinstance(v, (int, np.integer, np.floating)):
        return float(v)
    if v is undefined:
        return math.nan
    if v is None:
        return 0.0
    if isinstance(v, str):
        try:
            return float(v) if v.strip() else 0.0
        except ValueError:
            return math.nan
    return math.nan


def to_int32(v):
    x = to_number(v)
    if x != x or x in (math.inf, -math.inf):
        return 0
    n = int(x) & 0xFFFFFFFF
    return n - 0x100000000 if n >= 0x80000000 else n


def to_uint32(v):
    return to_int32(v) & 0xFFFFFFFF


def truthy(v):
    if isinstance(v, float):
        return not (v == 0.0 or v != v)
    if v is undefined or v is None:
        return False
    if isinstance(v, bool):
        return v
    if isinstance(v, str):
        return len(v) > 0
    return True


def js_round(x):
    """Math.round: nearest integer, ties toward +Infinity."""
    if x != x or x in (math.inf, -math.inf):
        return x
    r = math.floor(x)
    return float(r + 1) if x - r >= 0.5 else float(r)


def _prop_key(k):
    """canonical property key: integers as int, everything else as its string form"""
    if isinstance(k, float):
        if k == math.floor(k) and abs(k) < 2 ** 53:
            return int(k)
        return repr(k)
    if isinstance(k, (int, np.integer)):
        return int(k)
    return str(k)


# --------------------------------------------------------------------------------------
# tokenizer
# --------------------------------------------------------------------------------------
_TOKEN = re.compile(r"""
    (?P<ws>\s+|//[^\n]*|/\*.*?\*/)
  | (?P<num>0[xX][0-9a-fA-F]+|(?:\d+\.?\d*|\.\d+)(?:[eE][+-]?\d+)?)
  | (?P<id>[A-Za-z_$][A-Za-z0-9_$]*)
  | (?P<str>'(?:[^'\\]|\\.)*'|"(?:[^"\\]|\\.)*")
  | (?P<op>>>>=|\*\*=|===|!==|>>>|<<=|>>=|\*\*|\+\+|--|&&|\|\||==|!=|<=|>=|\+=|-=|\*=|/=|%=|&=|\|=|\^=|<<|>>|=>|[{}()\[\];,<>+\-*/%&|^!~?:=.])
""", re.X | re.S)

KEYWORDS = {"var", "let", "const", "function", "return", "if", "else", "for", "while", "break",
            "continue", "new", "this", "class", "extends", "super", "static", "get", "throw",
            "true", "false", "null", "undefined", "typeof"}


def tokenize(src):
    out, pos = [], 0
    while pos < len(src):
        m = _TOKEN.match(src, pos)
        if not m:
            raise SyntaxError(f"jsmini: cannot tokenize at {pos}: {src[pos:pos + 30]!r}")
        pos = m.end()
        kind = m.lastgroup
        if kind == "ws":
            continue
        out.append((kind, m.group(kind)))
    out.append(("eof", ""))
    return out


# --------------------------------------------------------------------------------------
# parser -> closures
# --------------------------------------------------------------------------------------
_BINARY_PREC = {
    "||": 1, "&&": 2, "|": 3, "^": 4, "&": 5,
    "==": 6, "!=": 6, "===": 6, "!==": 6,
    "<": 7, ">": 7, "<=": 7, ">=": 7,
    "<<": 8, ">>": 8, ">>>": 8,
    "+": 9, "-": 9, "*": 10, "/": 10, "%": 10, "**": 11,
}
_ASSIGN_OPS = {"=", "+=", "-=", "*=", "/=", "%=", "&=", "|=", "^=", "<<=", ">>=", ">>>=", "**="}


def _binary(op, a, b):
    if op == "+":
        if isinstance(a, str) or isinstance(b, str):
            return _to_string(a) + _to_string(b)
        return to_number(a) + to_number(b)
    if op == "-":
        return to_number(a) - to_number(b)
    if op == "*":
        return to_number(a) * to_number(b)
    if op == "/":
        x, y = to_number(a), to_number(b)
        if y == 0.0:
            if x != x or x == 0.0:
                return math.nan
            return math.copysign(math.inf, x) * math.copysign(1.0, y)
        return x / y
    if op == "%":
        x, y = to_number(a), to_number(b)
        if y == 0.0 or x in (math.inf, -math.inf) or x != x or y != y:
            return math.nan
        return math.fmod(x, y)
    if op == "**":
        try:
            return math.pow(to_number(a), to_number(b))
        except (OverflowError, ValueError):
            return math.nan
    if op in ("<", ">", "<=", ">="):
        x, y = to_number(a), to_number(b)
        return {"<": x < y, ">": x > y, "<=": x <= y, ">=": x >= y}[op]
    if op in ("==", "==="):
        return _equals(a, b)
    if op in ("!=", "!=="):
        return not _equals(a, b)
    if op == "&":
        return float(to_int32(to_int32(a) & to_int32(b)))
    if op == "|":
        return float(to_int32(to_int32(a) | to_int32(b)))
    if op == "^":
        return float(to_int32(to_int32(a) ^ to_int32(b)))
    if op == "<<":
        return float(to_int32((to_int32(a) << (to_uint32(b) & 31)) & 0xFFFFFFFF))
    if op == ">>":
        return float(to_int32(a) >> (to_uint32(b) & 31))
    if op == ">>>":
        return float(to_uint32(a) >> (to_uint32(b) & 31))
    raise NotImplementedError(op)


def _equals(a, b):
    if isinstance(a, (float, int)) and isinstance(b, (float, int)) and not isinstance(a, bool) and not isinstance(b, bool):
        return float(a) == float(b)
    if a is b:
        return True
    if isinstance(a, str) and isinstance(b, str):
        return a == b
    if isinstance(a, bool) and isinstance(b, bool):
        return a == b
    if (a is None or a is undefined) and (b is None or b is undefined):
        return True
    return False


def _to_string(v):
    if isinstance(v, str):
        return v
    if isinstance(v, float):
        return str(int(v)) if v == math.floor(v) and abs(v) < 1e21 else repr(v)
    if v is undefined:
        return "undefined"
    if v is None:
        return "null"
    if isinstance(v, bool):
        return "true" if v else "false"
    return "[object Object]"


def get_member(obj, key):
    if isinstance(obj, JSTyped):
        if isinstance(key, float):
            i = int(key)
            if key == i and 0 <= i < obj.a.shape[0]:
                return float(obj.a[i])
            return undefined
        if key == "length":
            return float(obj.a.shape[0])
        return _typed_method(obj, key)
    if isinstance(obj, JSArray):
        if isinstance(key, float):
            i = int(key)
            if key == i and i >= 0:
                return obj.items[i] if i < len(obj.items) else undefined
            return obj.props.get(_prop_key(key), undefined)
        if key == "length":
            return float(len(obj.items))
        if key == "fill":
            return JSFunction("fill", [], None, None, native=lambda this, a: _array_fill(this, a))
        return obj.props.get(_prop_key(key), undefined)
    if isinstance(obj, JSObject):
        return obj.get(key if isinstance(key, str) else _prop_key(key))
    if obj is undefined or obj is None:
        raise JSError(f"TypeError: cannot read properties of {obj!r} (reading {key!r})")
    return undefined


def set_member(obj, key, value):
    if isinstance(obj, JSTyped):
        if isinstance(key, float):
            i = int(key)
            if key == i and 0 <= i < obj.a.shape[0]:
                if obj.a.dtype == np.float32:
                    obj.a[i] = np.float32(to_number(value))      # one rounding, like a Float32Array store
                else:
                    obj.a[i] = to_int32(value)
            return                                             # out-of-range typed writes are ignored
        return
    if isinstance(obj, JSArray):
        if isinstance(key, float):
            i = int(key)
            if key == i and i >= 0:
                if i >= len(obj.items):
                    obj.items.extend([undefined] * (i + 1 - len(obj.items)))
                obj.items[i] = value
                return
        obj.props[_prop_key(key)] = value                     # e.g. arr[-2] = x : a named property
        return
    if isinstance(obj, JSObject):
        obj.set(key if isinstance(key, str) else _prop_key(key), value)
        return
    raise JSError(f"TypeError: cannot set properties of {obj!r}")


def _array_fill(this, args):
    v = args[0] if args else undefined
    n = len(this.items)
    start = int(to_number(args[1])) if len(args) > 1 and args[1] is not undefined else 0
    end = int(to_number(args[2])) if len(args) > 2 and args[2] is not undefined else n
    start = max(n + start, 0) if start < 0 else min(start, n)
    end = max(n + end, 0) if end < 0 else min(end, n)
    for i in range(start, end):
        this.items[i] = v
    return this


def _typed_method(obj, key):
    a = obj.a
    n = a.shape[0]

    def clamp(v, default):
        if v is undefined:
            return default
        i = int(to_number(v))
        return max(n + i, 0) if i < 0 else min(i, n)

    if key == "fill":
        def fill(this, args):
            start = clamp(args[1] if len(args) > 1 else undefined, 0)
            end = clamp(args[2] if len(args) > 2 else undefined, n)
            a[start:end] = to_number(args[0]) if a.dtype == np.float32 else to_int32(args[0])
            return obj
        return JSFunction("fill", [], None, None, native=fill)
    if key == "set":
        def set_(this, args):
            src = args[0]
            off = int(to_number(args[1])) if len(args) > 1 and args[1] is not undefined else 0
            data = src.a if isinstance(src, JSTyped) else np.array([to_number(x) for x in src.items])
            if off + data.shape[0] > n:
                raise JSError("RangeError: offset is out of bounds")
            a[off:off + data.shape[0]] = data
            return undefined
        return JSFunction("set", [], None, None, native=set_)
    if key == "subarray":
        def subarray(this, args):
            b = clamp(args[0] if args else undefined, 0)
            e = clamp(args[1] if len(args) > 1 else undefined, n)
            return JSTyped(a[b:max(b, e)])
        return JSFunction("subarray", [], None, None, native=subarray)
    if key == "copyWithin":
        def copy_within(this, args):
            target = clamp(args[0], 0)
            start = clamp(args[1] if len(args) > 1 else undefined, 0)
            end = clamp(args[2] if len(args) > 2 else undefined, n)
            count = min(end - start, n - target)
            if count > 0:
                a[target:target + count] = a[start:start + count].copy()
            return obj
        return JSFunction("copyWithin", [], None, None, native=copy_within)
    return obj.props.get(key, undefined)


class Parser:
    def __init__(self, src):
        self.toks = tokenize(src)
        self.i = 0
        self.fn_stack = []        # per function: {"decls": {name: depth}, "depth": n, "hoisted": set}

    # -- token helpers -----------------------------------------------------------------
    def peek(self, k=0):
        return self.toks[self.i + k]

    def at(self, text):
        t = self.toks[self.i]
        return t[1] == text and t[0] in ("op", "id")

    def eat(self, text):
        if self.at(text):
            self.i += 1
            return True
        return False

    def expect(self, text):
        if not self.eat(text):
            raise SyntaxError(f"jsmini: expected {text!r}, got {self.peek()[1]!r} (token {self.i})")

    def ident(self):
        kind, text = self.toks[self.i]
        if kind != "id":
            raise SyntaxError(f"jsmini: expected identifier, got {text!r}")
        self.i += 1
        return text

    # -- declarations bookkeeping ----------------------------------------------------------
    def declare(self, name, kind):
        f = self.fn_stack[-1]
        depth = f["depth"]
        if kind != "var" and name in f["decls"] and f["decls"][name] < depth:
            raise NotImplementedError(f"jsmini: {kind} {name} would shadow an outer binding")
        f["decls"].setdefault(name, depth)
        f["hoisted"].add(name)

    # -- program / statements ------------------------------------------------------------------
    def parse_program(self):
        self.fn_stack.append({"decls": {}, "depth": 0, "hoisted": set()})
        stmts = []
        while self.peek()[0] != "eof":
            stmts.append(self.statement())
        hoisted = tuple(self.fn_stack.pop()["hoisted"])
        return _block(stmts), hoisted

    def block(self):
        self.expect("{")
        self.fn_stack[-1]["depth"] += 1
        stmts = []
        while not self.at("}"):
            stmts.append(self.statement())
        self.expect("}")
        self.fn_stack[-1]["depth"] -= 1
        return _block(stmts)

    def statement(self):
        kind, text = self.peek()
        if kind == "str" and self.peek(1)[1] == ";":            # "use strict";
            self.i += 2
            return lambda s: None
        if text == "{" and kind == "op":
            return self.block()
        if text == ";" and kind == "op":
            self.i += 1
            return lambda s: None
        if kind == "id":
            if text in ("var", "let", "const"):
                st = self.var_decl()
                self.eat(";")
                return st
            if text == "function":
                self.i += 1
                name = self.ident()
                fn = self.function_rest(name)
                self.declare(name, "var")
                return lambda s, name=name, fn=fn: s.vars.__setitem__(name, fn(s))
            if text == "class":
                return self.class_decl()
            if text == "return":
                self.i += 1
                if self.at(";") or self.at("}"):
                    self.eat(";")
                    def ret0(s):
                        raise _Return(undefined)
                    return ret0
                e = self.expression()
                self.eat(";")
                def ret(s, e=e):
                    raise _Return(e(s))
                return ret
            if text == "if":
                self.i += 1
                self.expect("(")
                c = self.expression()
                self.expect(")")
                a = self.statement()
                b = self.statement() if self.eat("else") else None
                if b is None:
                    def if1(s, c=c, a=a):
                        if truthy(c(s)):
                            a(s)
                    return if1
                def if2(s, c=c, a=a, b=b):
                    if truthy(c(s)):
                        a(s)
                    else:
                        b(s)
                return if2
            if text == "for":
                return self.for_stmt()
            if text == "while":
                self.i += 1
                self.expect("(")
                c = self.expression()
                self.expect(")")
                body = self.statement()
                def wh(s, c=c, body=body):
                    while truthy(c(s)):
                        try:
                            body(s)
                        except _Continue:
                            continue
                        except _Break:
                            break
                return wh
            if text == "break":
                self.i += 1
                self.eat(";")
                def br(s):
                    raise _Break()
                return br
            if text == "continue":
                self.i += 1
                self.eat(";")
                def co(s):
                    raise _Continue()
                return co
            if text == "throw":
                self.i += 1
                e = self.expression()
                self.eat(";")
                def th(s, e=e):
                    raise JSError(e(s))
                return th
        e = self.expression()
        self.eat(";")
        return e

    def var_decl(self):
        kind = self.ident()
        decls = []
        while True:
            name = self.ident()
            self.declare(name, kind)
            init = self.assignment() if self.eat("=") else None
            decls.append((name, init))
            if not self.eat(","):
                break
        def run(s, decls=decls, kind=kind):
            for name, init in decls:
                if init is not None:
                    s.lookup_or_self(name).vars[name] = init(s)
                elif kind != "var":
                    s.lookup_or_self(name).vars[name] = undefined
        return run

    def for_stmt(self):
        self.i += 1
        self.expect("(")
        self.fn_stack[-1]["depth"] += 1
        init = None
        if not self.at(";"):
            init = self.var_decl() if self.peek()[1] in ("var", "let", "const") else self.expression()
        self.expect(";")
        test = None if self.at(";") else self.expression()
        self.expect(";")
        update = None if self.at(")") else self.expression()
        self.expect(")")
        body = self.statement()
        self.fn_stack[-1]["depth"] -= 1
        def run(s, init=init, test=test, update=update, body=body):
            if init is not None:
                init(s)
            while test is None or truthy(test(s)):
                try:
                    body(s)
                except _Continue:
                    pass
                except _Break:
                    break
                if update is not None:
                    update(s)
        return run

    def function_rest(self, name):
        """after the (optional) name: (params) { body } -> maker(scope) -> JSFunction"""
        self.expect("(")
        params = []
        while not self.at(")"):
            params.append(self.ident())
            if not self.eat(","):
                break
        self.expect(")")
        self.fn_stack.append({"decls": {p: 0 for p in params}, "depth": 0, "hoisted": set()})
        body = self.block()
        hoisted = tuple(self.fn_stack.pop()["hoisted"])
        def make(s, name=name, params=params, body=body, hoisted=hoisted):
            f = JSFunction(name, params, body, s)
            f.hoisted = hoisted
            return f
        return make

    def class_decl(self):
        self.i += 1
        name = self.ident()
        parent = self.expression_member_only() if self.eat("extends") else None
        self.expect("{")
        members = []          # (kind, name, maker) kind in method/static/static_get/get
        while not self.at("}"):
            static = False
            getter = False
            if self.at("static"):
                self.i += 1
                static = True
            if self.at("get") and self.peek(1)[0] == "id":
                self.i += 1
                getter = True
            mname = self.ident()
            maker = self.function_rest(mname)
            members.append((static, getter, mname, maker))
        self.expect("}")
        self.declare(name, "let")

        def run(s, name=name, parent=parent, members=members):
            par = parent(s) if parent is not None else None
            ctor = None
            for static, getter, mname, maker in members:
                if mname == "constructor":
                    ctor = maker(s)
            if ctor is None:
                if par is not None:
                    ctor = JSFunction(name, [], None, s, native=None)
                    ctor.native = lambda this, args, par=par: par.call(this, args)
                else:
                    ctor = JSFunction(name, [], None, s, native=lambda this, args: undefined)
            ctor.name, ctor.is_class, ctor.parent = name, True, par
            proto = ctor.props["prototype"]
            if par is not None:
                proto.proto = par.props["prototype"]
                ctor.proto = par
            for static, getter, mname, maker in members:
                if mname == "constructor":
                    continue
                fn = maker(s)
                fn.parent = par                   # for super() resolution inside methods (unused here)
                target = ctor if static else proto
                if getter:
                    if target.getters is None:
                        target.getters = {}
                    target.getters[mname] = fn
                else:
                    target.props[mname] = fn
            s.lookup_or_self(name).vars[name] = ctor
        return run

    def expression_member_only(self):
        return self.unary()

    # -- expressions -------------------------------------------------------------------------
    def expression(self):
        e = self.assignment()
        if self.at(","):
            parts = [e]
            while self.eat(","):
                parts.append(self.assignment())
            def seq(s, parts=parts):
                v = undefined
                for p in parts:
                    v = p(s)
                return v
            return seq
        return e

    def assignment(self):
        start = self.i
        left = self.conditional()
        kind, text = self.peek()
        if kind == "op" and text in _ASSIGN_OPS:
            ref = getattr(left, "ref", None)
            if ref is None:
                raise SyntaxError(f"jsmini: invalid assignment target near token {start}")
            self.i += 1
            right = self.assignment()
            setter = ref[1]
            if text == "=":
                def assign(s, setter=setter, right=right):
                    v = right(s)
                    setter(s, v)
                    return v
                return assign
            op = text[:-1]
            def compound(s, ref=ref, right=right, op=op):
                target = ref[2](s) if len(ref) > 2 else None
                old = ref[3](s, target) if len(ref) > 2 else ref[0](s)
                v = _binary(op, old, right(s))
                if len(ref) > 2:
                    ref[4](s, target, v)
                else:
                    ref[1](s, v)
                return v
            return compound
        return left

    def conditional(self):
        c = self.binary(0)
        if self.eat("?"):
            a = self.assignment()
            self.expect(":")
            b = self.assignment()
            return lambda s, c=c, a=a, b=b: a(s) if truthy(c(s)) else b(s)
        return c

    def binary(self, min_prec):
        left = self.unary()
        while True:
            kind, op = self.peek()
            prec = _BINARY_PREC.get(op) if kind == "op" else None
            if prec is None or prec < min_prec:
                return left
            self.i += 1
            right = self.binary(prec if op == "**" else prec + 1)
            left = self._make_binary(op, left, right)

    @staticmethod
    def _make_binary(op, a, b):
        if op == "&&":
            def land(s, a=a, b=b):
                v = a(s)
                return b(s) if truthy(v) else v
            return land
        if op == "||":
            def lor(s, a=a, b=b):
                v = a(s)
                return v if truthy(v) else b(s)
            return lor
        if op in ("+", "-", "*"):
            # fast paths for the common float / float case
            if op == "+":
                def add(s, a=a, b=b):
                    x, y = a(s), b(s)
                    if type(x) is float and type(y) is float:
                        return x + y
                    return _binary("+", x, y)
                return add
            if op == "-":
                def sub(s, a=a, b=b):
                    x, y = a(s), b(s)
                    if type(x) is float and type(y) is float:
                        return x - y
                    return _binary("-", x, y)
                return sub
            def mul(s, a=a, b=b):
                x, y = a(s), b(s)
                if type(x) is float and type(y) is float:
                    return x * y
                return _binary("*", x, y)
            return mul
        return lambda s, a=a, b=b, op=op: _binary(op, a(s), b(s))

    def unary(self):
        kind, text = self.peek()
        if kind == "op" and text in ("!", "-", "+", "~"):
            self.i += 1
            e = self.unary()
            if text == "!":
                return lambda s, e=e: not truthy(e(s))
            if text == "-":
                return lambda s, e=e: -to_number(e(s))
            if text == "+":
                return lambda s, e=e: to_number(e(s))
            return lambda s, e=e: float(to_int32(~to_int32(e(s))))
        if kind == "op" and text in ("++", "--"):
            self.i += 1
            e = self.unary()
            return self._update(e, text, prefix=True)
        if kind == "id" and text == "typeof":
            self.i += 1
            e = self.unary()
            def typeof(s, e=e):
                try:
                    v = e(s)
                except JSError:
                    return "undefined"
                if isinstance(v, JSFunction):
                    return "function"
                if isinstance(v, float):
                    return "number"
                if isinstance(v, str):
                    return "string"
                if isinstance(v, bool):
                    return "boolean"
                if v is undefined:
                    return "undefined"
                return "object"
            return typeof
        e = self.postfix()
        # ** binds tighter than unary on its left operand only through binary(); nothing to do
        return e

    def _update(self, e, op, prefix):
        ref = getattr(e, "ref", None)
        if ref is None:
            raise SyntaxError("jsmini: invalid update target")
        delta = 1.0 if op == "++" else -1.0
        def upd(s, ref=ref, delta=delta, prefix=prefix):
            if len(ref) > 2:
                target = ref[2](s)
                old = to_number(ref[3](s, target))
                ref[4](s, target, old + delta)
            else:
                old = to_number(ref[0](s))
                ref[1](s, old + delta)
            return old + delta if prefix else old
        return upd

    def postfix(self):
        e = self.call_member()
        kind, text = self.peek()
        if kind == "op" and text in ("++", "--"):
            self.i += 1
            return self._update(e, text, prefix=False)
        return e

    def arguments(self):
        self.expect("(")
        args = []
        while not self.at(")"):
            args.append(self.assignment())
            if not self.eat(","):
                break
        self.expect(")")
        return args

    def call_member(self):
        kind, text = self.peek()
        if kind == "id" and text == "new":
            self.i += 1
            callee = self.member_only()
            args = self.arguments() if self.at("(") else []
            def new(s, callee=callee, args=args):
                f = callee(s)
                if not isinstance(f, JSFunction):
                    raise JSError("TypeError: not a constructor")
                return f.construct([a(s) for a in args])
            e = new
        elif kind == "id" and text == "super":
            self.i += 1
            args = self.arguments()
            def sup(s, args=args):
                fscope = s
                while fscope is not None and fscope.fn is None:
                    fscope = fscope.parent
                parent = fscope.fn.parent
                this = s.lookup("this").vars["this"]
                return parent.call(this, [a(s) for a in args])
            e = sup
        else:
            e = self.primary()
        return self.member_suffixes(e, allow_call=True)

    def member_only(self):
        return self.member_suffixes(self.primary(), allow_call=False)

    def member_suffixes(self, e, allow_call):
        while True:
            if self.at("."):
                self.i += 1
                name = self.ident()
                e = self._member(e, lambda s, name=name: name)
            elif self.at("["):
                self.i += 1
                k = self.expression()
                self.expect("]")
                e = self._member(e, k)
            elif allow_call and self.at("("):
                args = self.arguments()
                e = self._call(e, args)
            else:
                return e

    @staticmethod
    def _member(obj, key):
        def get(s, obj=obj, key=key):
            return get_member(obj(s), key(s))
        # ref protocol for members: (getter, setter, eval_target, get_from_target, set_on_target)
        def setter(s, v, obj=obj, key=key):
            set_member(obj(s), key(s), v)
        def target(s, obj=obj, key=key):
            return (obj(s), key(s))
        def tget(s, t):
            return get_member(t[0], t[1])
        def tset(s, t, v):
            set_member(t[0], t[1], v)
        get.ref = (get, setter, target, tget, tset)
        get.obj, get.key = obj, key
        return get

    @staticmethod
    def _call(callee, args):
        obj = getattr(callee, "obj", None)
        if obj is not None:                           # method call: this = the object
            key = callee.key
            def mcall(s, obj=obj, key=key, args=args):
                o = obj(s)
                f = get_member(o, key(s))
                if not isinstance(f, JSFunction):
                    raise JSError(f"TypeError: {key(s)!r} is not a function")
                return f.call(o, [a(s) for a in args])
            return mcall
        def call(s, callee=callee, args=args):
            f = callee(s)
            if not isinstance(f, JSFunction):
                raise JSError("TypeError: not a function")
            return f.call(undefined, [a(s) for a in args])
        return call

    def primary(self):
        kind, text = self.peek()
        self.i += 1
        if kind == "num":
            v = float(int(text, 16)) if text[:2] in ("0x", "0X") else float(text)
            return lambda s, v=v: v
        if kind == "str":
            v = bytes(text[1:-1], "utf-8").decode("unicode_escape")
            return lambda s, v=v: v
        if kind == "op" and text == "(":
            e = self.expression()
            self.expect(")")
            return e
        if kind == "op" and text == "[":
            items = []
            while not self.at("]"):
                items.append(self.assignment())
                if not self.eat(","):
                    break
            self.expect("]")
            return lambda s, items=items: JSArray([i(s) for i in items])
        if kind == "op" and text == "{":
            props = []
            while not self.at("}"):
                k = self.peek()
                self.i += 1
                key = k[1][1:-1] if k[0] == "str" else k[1]
                self.expect(":")
                props.append((key, self.assignment()))
                if not self.eat(","):
                    break
            self.expect("}")
            def obj(s, props=props):
                o = JSObject(None)
                for k, v in props:
                    o.props[k] = v(s)
                return o
            return obj
        if kind == "id":
            if text == "function":
                name = self.ident() if self.peek()[0] == "id" else ""
                return self.function_rest(name)
            if text == "true":
                return lambda s: True
            if text == "false":
                return lambda s: False
            if text == "null":
                return lambda s: None
            if text == "undefined":
                return lambda s: undefined
            name = text
            def var(s, name=name):
                sc = s
                while sc is not None:
                    vs = sc.vars
                    if name in vs:
                        return vs[name]
                    sc = sc.parent
                raise JSError(f"ReferenceError: {name} is not defined")
            def setvar(s, v, name=name):
                sc = s.lookup(name)
                if sc is None:
                    raise JSError(f"ReferenceError: {name} is not defined")
                sc.vars[name] = v
            var.ref = (var, setvar)
            return var
        raise SyntaxError(f"jsmini: unexpected token {text!r} (token {self.i - 1})")


def _scope_lookup_or_self(self, name):
    return self.lookup(name) or self


Scope.lookup_or_self = _scope_lookup_or_self


def _block(stmts):
    if len(stmts) == 1:
        return stmts[0]
    def run(s, stmts=stmts):
        for st in stmts:
            st(s)
    return run


# --------------------------------------------------------------------------------------
# runtime
# --------------------------------------------------------------------------------------

def _native(fn):
    return JSFunction(getattr(fn, "__name__", "native"), [], None, None, native=fn)


def make_globals(const_overrides=None):
    g = Scope(None)
    m = JSObject(None)
    m.props.update({
        "PI": math.pi,
        "cos": _native(lambda this, a: math.cos(to_number(a[0]))),
        "sin": _native(lambda this, a: math.sin(to_number(a[0]))),
        "round": _native(lambda this, a: js_round(to_number(a[0]))),
        "floor": _native(lambda this, a: float(math.floor(to_number(a[0]))) if math.isfinite(to_number(a[0])) else to_number(a[0])),
        "ceil": _native(lambda this, a: float(math.ceil(to_number(a[0]))) if math.isfinite(to_number(a[0])) else to_number(a[0])),
        "sqrt": _native(lambda this, a: math.sqrt(to_number(a[0])) if to_number(a[0]) >= 0 else math.nan),
        "abs": _native(lambda this, a: abs(to_number(a[0]))),
        "max": _native(lambda this, a: max(to_number(x) for x in a)),
        "min": _native(lambda this, a: min(to_number(x) for x in a)),
    })
    g.vars["Math"] = m

    def array_ctor(this, args):
        if len(args) == 1 and isinstance(args[0], float):
            return JSArray([undefined] * int(args[0]))
        return JSArray(list(args))
    arr = _native(array_ctor)
    arr.ctor_native = lambda args: array_ctor(None, args)
    g.vars["Array"] = arr

    def typed(dtype):
        def ctor(this, args):
            if args and isinstance(args[0], (JSArray, JSTyped)):
                src = args[0]
                data = src.a if isinstance(src, JSTyped) else [to_number(x) for x in src.items]
                return JSTyped(np.array(data, dtype=dtype))
            return JSTyped(np.zeros(int(to_number(args[0])) if args else 0, dtype=dtype))
        f = _native(ctor)
        f.ctor_native = lambda args: ctor(None, args)
        return f
    g.vars["Float32Array"] = typed(np.float32)
    g.vars["Int32Array"] = typed(np.int32)

    def error_ctor(this, args):
        o = JSObject(None)
        o.props["message"] = args[0] if args else ""
        return o
    err = _native(error_ctor)
    err.ctor_native = lambda args: error_ctor(None, args)
    g.vars["Error"] = err

    console = JSObject(None)
    def cassert(this, args):
        if not truthy(args[0] if args else undefined):
            raise JSError("console.assert failed: " + _to_string(args[1] if len(args) > 1 else ""))
        return undefined
    console.props["assert"] = _native(cassert)
    console.props["log"] = _native(lambda this, a: undefined)
    g.vars["console"] = console
    g.vars["undefined"] = undefined
    g.vars["NaN"] = math.nan
    g.vars["Infinity"] = math.inf
    return g


def run_module(src, globals_scope, require=None, const_overrides=None):
    """Run `src` as a CommonJS module body; returns module.exports.

    const_overrides {NAME: number}: after the module's top-level code has run a constant's
    declaration, its value is replaced.  Used ONLY for the two size constants the reference
    hard-codes (BUFFERED_BLOCK_SIZE, WEBAUDIO_BLOCK_SIZE) so that other frame / hop sizes can be
    exercised with the otherwise unmodified source."""
    parser = Parser(src)
    body, hoisted = parser.parse_program()
    scope = Scope(globals_scope)
    for name in hoisted:
        scope.vars[name] = undefined
    module = JSObject(None)
    module.props["exports"] = JSObject(None)
    scope.vars["module"] = module
    scope.vars["exports"] = module.props["exports"]
    scope.vars["this"] = undefined
    if require is not None:
        scope.vars["require"] = _native(lambda this, a: require(a[0]))
    if const_overrides:
        class _Pinned(dict):
            def __setitem__(self, k, v, ov=const_overrides):
                dict.__setitem__(self, k, float(ov[k]) if k in ov and v is not undefined else v)
        pinned = _Pinned(scope.vars)
        scope.vars = pinned
    body(scope)
    return module.props["exports"], scope


# --------------------------------------------------------------------------------------
# the reference, loaded from /root/reference
# --------------------------------------------------------------------------------------
REFERENCE_ROOT = "/root/reference"


def load_reference(frame_size=None, hop_size=None, root=REFERENCE_ROOT):
    """Load src/ola-processor.js + src/phase-vocoder.js + fft.js exactly as they are on disk
    and return (PhaseVocoderProcessor class, globals scope).

    frame_size / hop_size override the two hard-coded constants (phase-vocoder.js:6,
    ola-processor.js:3); None keeps the reference's own 2048 / 128."""
    import os

    with open(os.path.join(root, "www", "phase-vocoder.js")) as f:
        bundle_lines = f.read().split("\n")
    # module 1 of the browserify bundle == fft.js 4.0.3 lib/fft.js (bundle lines 2..509)
    assert bundle_lines[3].startswith("function FFT(size)"), "unexpected bundle layout"
    end = next(i for i, line in enumerate(bundle_lines) if line.startswith("},{}],2:[function"))
    fft_src = "\n".join(bundle_lines[1:end])
    with open(os.path.join(root, "src", "ola-processor.js")) as f:
        ola_src = f.read()
    with open(os.path.join(root, "src", "phase-vocoder.js")) as f:
        pv_src = f.read()

    g = make_globals()
    registry = {}
    awp = JSFunction("AudioWorkletProcessor", [], None, None, native=lambda this, args: undefined)
    awp.is_class = True
    g.vars["AudioWorkletProcessor"] = awp
    g.vars["registerProcessor"] = _native(lambda this, a: registry.__setitem__(a[0], a[1]) or undefined)

    cache = {}

    def require(name):
        if name not in cache:
            if name == "fft.js":
                cache[name] = run_module(fft_src, g)[0]
            elif name == "./ola-processor.js":
                ov = {"WEBAUDIO_BLOCK_SIZE": hop_size} if hop_size else None
                cache[name] = run_module(ola_src, g, require, ov)[0]
            else:
                raise JSError(f"Cannot find module '{name}'")
        return cache[name]

    ov = {"BUFFERED_BLOCK_SIZE": frame_size} if frame_size else None
    run_module(pv_src, g, require, ov)
    return registry["phase-vocoder-processor"], g


class ReferenceProcessor:
    """The reference's PhaseVocoderProcessor instance, driven from Python."""

    def __init__(self, frame_size=None, hop_size=None, num_inputs=1, num_outputs=1):
        cls, self.globals = load_reference(frame_size, hop_size)
        options = JSObject(None)
        options.props["numberOfInputs"] = float(num_inputs)
        options.props["numberOfOutputs"] = float(num_outputs)
        self.obj = cls.construct([options])
        self.hop = int(self.obj.get("hopSize"))
        self.frame = int(self.obj.get("blockSize"))

    def process(self, inputs, outputs, pitch_factor):
        """inputs / outputs: list (per input) of list (per channel) of float32 numpy arrays of
        one quantum; outputs are filled in place.  Mirrors process(inputs, outputs, params)."""
        def wrap(nested):
            return JSArray([JSArray([JSTyped(ch) for ch in inp]) for inp in nested])
        params = JSObject(None)
        params.props["pitchFactor"] = JSTyped(np.array([pitch_factor], dtype=np.float32))
        fn = self.obj.get("process")
        return fn.call(self.obj, [wrap(inputs), wrap(outputs), params])

    def run(self, signal, pitch_factor):
        """signal [C][T*hop] float32 (one input, C channels) -> output of T process() calls."""
        signal = np.ascontiguousarray(signal, np.float32)
        C, total = signal.shape
        out = np.zeros_like(signal)
        for t in range(total // self.hop):
            sl = slice(t * self.hop, (t + 1) * self.hop)
            ins = [[signal[c, sl].copy() for c in range(C)]]
            outs = [[np.zeros(self.hop, np.float32) for _ in range(C)]]
            self.process(ins, outs, pitch_factor)
            for c in range(C):
                out[c, sl] = outs[0][c]
        return out

    @property
    def time_cursor(self):
        return self.obj.get("timeCursor")
