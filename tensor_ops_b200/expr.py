"""Symbolic scalar type standing in for `ElemT t` on the device backend.

The reference's elementwise methods take HOST closures — `liftT :: (Vec n (ElemT t) -> ElemT t) -> ...`
(src/TensorOps/Types.hs:56-59), `liftB` (src/TensorOps/BLAS.hs:92-96) — which a GPU cannot call per element.
The closure is therefore applied ONCE to symbolic variables; the resulting expression tree is differentiated
symbolically where the reference uses `ad`'s `diff` / `grad` (src/TensorOps/TOp.hs:212,246) and serialised to the
postfix bytecode `tops_lift` executes on the device (include/tops_b200.h, TOPS_OP_*).

`Expr` implements the Num / Fractional / Floating surface the reference's activation and loss functions use
(src/TensorOps/Learn/NeuralNet.hs:42-77).
"""
from __future__ import annotations

import math
from typing import Callable, List, Sequence, Tuple, Union

from ._lib import OPCODES

Number = Union[int, float]

_BINARY = {"add": "ADD", "sub": "SUB", "mul": "MUL", "div": "DIV", "max": "MAX", "min": "MIN", "pow": "POW"}
_UNARY = {"neg": "NEG", "exp": "EXP", "log": "LOG", "recip": "RECIP", "sqrt": "SQRT", "tanh": "TANH", "abs": "ABS",
          "signum": "SIGNUM", "logistic": "LOGISTIC", "sin": "SIN", "cos": "COS"}


class Expr:
    __slots__ = ("op", "args", "value")

    def __init__(self, op: str, args: Tuple["Expr", ...] = (), value=None):
        self.op, self.args, self.value = op, args, value

    # -- constructors
    @staticmethod
    def var(i: int) -> "Expr":
        return Expr("var", (), int(i))

    @staticmethod
    def const(v: Number) -> "Expr":
        return Expr("const", (), float(v))

    @staticmethod
    def lift(x) -> "Expr":
        return x if isinstance(x, Expr) else Expr.const(x)

    def is_const(self, v=None) -> bool:
        return self.op == "const" and (v is None or self.value == v)

    # -- Num / Fractional
    def __add__(self, o): return _bin("add", self, Expr.lift(o))
    def __radd__(self, o): return _bin("add", Expr.lift(o), self)
    def __sub__(self, o): return _bin("sub", self, Expr.lift(o))
    def __rsub__(self, o): return _bin("sub", Expr.lift(o), self)
    def __mul__(self, o): return _bin("mul", self, Expr.lift(o))
    def __rmul__(self, o): return _bin("mul", Expr.lift(o), self)
    def __truediv__(self, o): return _bin("div", self, Expr.lift(o))
    def __rtruediv__(self, o): return _bin("div", Expr.lift(o), self)
    def __pow__(self, o): return _bin("pow", self, Expr.lift(o))
    def __neg__(self): return _un("neg", self)
    def __abs__(self): return _un("abs", self)

    def __repr__(self):
        if self.op == "var": return f"x{self.value}"
        if self.op == "const": return repr(self.value)
        return f"{self.op}({', '.join(map(repr, self.args))})"

    def key(self):
        return (self.op, self.value, tuple(a.key() for a in self.args))


def _bin(op: str, a: Expr, b: Expr) -> Expr:
    if a.op == "const" and b.op == "const":
        x, y = a.value, b.value
        try:
            v = {"add": x + y, "sub": x - y, "mul": x * y, "div": x / y if y != 0 else None,
                 "max": max(x, y), "min": min(x, y), "pow": x ** y}[op]
        except (OverflowError, ValueError, ZeroDivisionError):
            v = None
        if isinstance(v, complex) or (isinstance(v, float) and v != v):
            v = None          # e.g. (-8.0) ** 0.5: Python answers with a complex; leave it to the device (NaN, like powf)
        if v is not None:
            return Expr.const(v)
    if op == "add":
        if a.is_const(0.0): return b
        if b.is_const(0.0): return a
    if op == "sub" and b.is_const(0.0): return a
    if op == "mul":
        if a.is_const(1.0): return b
        if b.is_const(1.0): return a
        if a.is_const(0.0) or b.is_const(0.0): return Expr.const(0.0)
    if op == "div" and b.is_const(1.0): return a
    return Expr(op, (a, b))


def _un(op: str, a: Expr) -> Expr:
    if a.op == "const":
        x = a.value
        try:
            v = {"neg": -x, "exp": math.exp(x), "log": math.log(x) if x > 0 else None, "recip": 1.0 / x if x != 0 else None,
                 "sqrt": math.sqrt(x) if x >= 0 else None, "tanh": math.tanh(x), "abs": abs(x),
                 "signum": float((x > 0) - (x < 0)), "logistic": 1.0 / (1.0 + math.exp(-x)), "sin": math.sin(x), "cos": math.cos(x)}[op]
        except (OverflowError, ValueError):
            v = None
        if v is not None:
            return Expr.const(v)
    if op == "neg" and a.op == "neg":
        return a.args[0]
    return Expr(op, (a,))


# -- the Floating surface, usable on Expr and on plain floats alike ----------------------------------------
def _dispatch(op: str, pyfn):
    def f(x):
        return _un(op, x) if isinstance(x, Expr) else pyfn(x)
    f.__name__ = op
    return f


exp = _dispatch("exp", math.exp)
log = _dispatch("log", math.log)
sqrt = _dispatch("sqrt", math.sqrt)
tanh = _dispatch("tanh", math.tanh)
sin = _dispatch("sin", math.sin)
cos = _dispatch("cos", math.cos)
recip = _dispatch("recip", lambda v: 1.0 / v)
signum = _dispatch("signum", lambda v: float((v > 0) - (v < 0)))


def maximum(a, b): return _bin("max", Expr.lift(a), Expr.lift(b))
def minimum(a, b): return _bin("min", Expr.lift(a), Expr.lift(b))


def logistic(x):
    """NeuralNet.hs:42-44."""
    return 1 / (1 + exp(-x))


def logistic_(x):
    """NeuralNet.hs:46-50: logix * (1 - logix), logix recomputed from x."""
    s = logistic(x)
    return s * (1 - s)


# -- symbolic differentiation (the role `ad` plays in TOp.hs:208-213,240-247) ---------------------------------
def diff(e: Expr, i: int) -> Expr:
    """d e / d x_i."""
    op, a = e.op, e.args
    if op == "var": return Expr.const(1.0 if e.value == i else 0.0)
    if op == "const": return Expr.const(0.0)
    if op == "add": return diff(a[0], i) + diff(a[1], i)
    if op == "sub": return diff(a[0], i) - diff(a[1], i)
    if op == "mul": return diff(a[0], i) * a[1] + a[0] * diff(a[1], i)
    if op == "div": return (diff(a[0], i) * a[1] - a[0] * diff(a[1], i)) / (a[1] * a[1])
    if op == "neg": return -diff(a[0], i)
    if op == "exp": return diff(a[0], i) * e
    if op == "log": return diff(a[0], i) / a[0]
    if op == "recip": return -diff(a[0], i) / (a[0] * a[0])
    if op == "sqrt": return diff(a[0], i) / (2 * e)
    if op == "tanh": return diff(a[0], i) * (1 - e * e)
    if op == "sin": return diff(a[0], i) * _un("cos", a[0])
    if op == "cos": return -(diff(a[0], i) * _un("sin", a[0]))
    if op == "abs": return diff(a[0], i) * _un("signum", a[0])
    if op == "signum": return Expr.const(0.0)
    if op == "logistic": return diff(a[0], i) * (e * (1 - e))
    if op == "pow":
        if a[1].op == "const":
            return diff(a[0], i) * (a[1] * _bin("pow", a[0], Expr.const(a[1].value - 1)))
        return e * (diff(a[1], i) * _un("log", a[0]) + a[1] * diff(a[0], i) / a[0])
    if op in ("max", "min"):
        raise NotImplementedError("max/min are not differentiable symbolically here")
    raise ValueError(op)


def trace(f: Callable[..., Expr], n: int) -> Expr:
    """Apply a host closure to n symbolic variables (what `liftT` receives: `Vec n (ElemT t) -> ElemT t`)."""
    return Expr.lift(f(*[Expr.var(k) for k in range(n)]))


# -- peephole: recognise the logistic so it runs as one opcode ------------------------------------------------
def _fuse(e: Expr) -> Expr:
    if e.op in ("var", "const"):
        return e
    args = tuple(_fuse(a) for a in e.args)
    e = Expr(e.op, args, e.value)
    # 1 / (1 + exp(-x))   or   recip(1 + exp(-x))
    den = None
    if e.op == "div" and args[0].is_const(1.0): den = args[1]
    if e.op == "recip": den = args[0]
    if den is not None and den.op == "add":
        for one, ex in ((den.args[0], den.args[1]), (den.args[1], den.args[0])):
            if one.is_const(1.0) and ex.op == "exp" and ex.args[0].op == "neg":
                return Expr("logistic", (ex.args[0].args[0],))
    return e


def compile_expr(e: Expr) -> Tuple[List[int], List[float]]:
    """Postfix bytecode for `tops_lift`: each instruction is (opcode << 16) | arg."""
    e = _fuse(e)
    code: List[int] = []
    consts: List[float] = []

    def emit(node: Expr):
        if node.op == "var":
            code.append((OPCODES["VAR"] << 16) | node.value)
        elif node.op == "const":
            v = float(node.value)
            if v not in consts:
                consts.append(v)
            code.append((OPCODES["CONST"] << 16) | consts.index(v))
        else:
            for a in node.args:
                emit(a)
            name = _BINARY.get(node.op) or _UNARY[node.op]
            code.append(OPCODES[name] << 16)

    emit(e)
    if len(code) > 64 or len(consts) > 16:
        raise ValueError(f"lifted expression too large for the device interpreter ({len(code)} ops, {len(consts)} constants)")
    return code, consts


def evaluate(e: Expr, xs: Sequence):
    """Reference evaluation of an expression on Python floats / NumPy arrays (host-side checks only)."""
    import numpy as np
    op, a = e.op, e.args
    if op == "var": return xs[e.value]
    if op == "const": return e.value
    v = [evaluate(t, xs) for t in a]
    if op == "add": return v[0] + v[1]
    if op == "sub": return v[0] - v[1]
    if op == "mul": return v[0] * v[1]
    if op == "div": return v[0] / v[1]
    if op == "pow": return v[0] ** v[1]
    if op == "max": return np.maximum(v[0], v[1])
    if op == "min": return np.minimum(v[0], v[1])
    if op == "neg": return -v[0]
    if op == "recip": return 1.0 / v[0]
    if op == "logistic": return 1.0 / (1.0 + np.exp(-v[0]))
    if op == "signum": return np.sign(v[0])
    return getattr(np, {"abs": "abs"}.get(op, op))(v[0])
