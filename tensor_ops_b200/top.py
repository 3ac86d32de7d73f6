"""The `TOp` algebra of tensor-ops (src/TensorOps/Types.hs:122-264, src/TensorOps/TOp.hs) over any backend `T`.

A `TOp` is a pair of closures — forward and vector-Jacobian product — polymorphic in the tensor type, exactly as
`data TOp` is rank-2 polymorphic over `Tensor t` (Types.hs:122-125).  Python has no type-class dispatch, so the
"dictionary" is passed explicitly: closures receive `T`, a class providing the `Tensor` methods as static methods
(`CuTensor` for single samples, `BatchT` for a whole batch).  `tag` carries an optional structural description that
lets the batched evaluator recognise fusable pipelines (ffLayer chains) without inspecting closures.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Callable, List, Optional, Sequence, Tuple

from . import expr as E

_zip = zip   # the reference's `zip` TOp shadows the builtin below

Prod = List[Any]


@dataclass
class TOp:
    run: Callable[[Any, Prod], Prod]            # T -> Prod t ns -> Prod t ms
    grad_: Callable[[Any, Prod, Prod], Prod]    # T -> Prod t ns -> Prod t ms -> Prod t ns      (gradTOp')
    n_in: int
    n_out: int
    tag: Tuple = ("opaque",)

    def __rshift__(self, other: "TOp") -> "TOp":   # (>>>)
        return compose(other, self)


def _T(xs: Prod, T=None):
    if T is not None:
        return T
    if not xs:
        raise ValueError("a TOp without inputs needs the backend passed explicitly (T=...)")
    return type(xs[0])


def runTOp(o: TOp, xs: Prod, T=None) -> Prod:
    """`runTOp` (Types.hs:123)."""
    return o.run(_T(xs, T), list(xs))


# Saved activations.  The reference's chain rule `g3 xs ds = g1 xs (g2 (f1 xs) ds)` (Types.hs:155) re-runs the forward of every
# prefix inside the gradient: a composition of n ops evaluates O(n^2) forwards (every `>>>` nests another `f1 xs`).  Tensors are
# immutable values, so within ONE gradient evaluation the forward of a composed prefix on the same input objects is the same
# value: `compose` looks it up here instead of launching it again — the reverse sweep then costs n forwards + n VJPs, the
# saved-activations schedule, with results identical to the reference's (same kernels on the same inputs, just not repeated).
_saved: Optional[dict] = None
forward_evals = 0          # instrumentation for the tests: forwards of composed prefixes actually executed


class saved_activations:
    """Scope in which forwards of composed prefixes are memoised (entered by gradTOp / gradTOp_; re-entrant)."""

    def __enter__(self):
        global _saved
        self.outer = _saved
        if _saved is None:
            _saved = {}
        return self

    def __exit__(self, *exc):
        global _saved
        _saved = self.outer
        return False


def _run_saved(o: "TOp", T, xs: Prod) -> Prod:
    global forward_evals
    if _saved is None:
        forward_evals += 1
        return o.run(T, xs)
    key = (id(o), id(T)) + tuple(id(x) for x in xs)
    hit = _saved.get(key)
    if hit is None:
        forward_evals += 1
        hit = (o.run(T, xs), o, list(xs))      # o and xs are kept alive so that their ids cannot be recycled inside the scope
        _saved[key] = hit
    return list(hit[0])


def gradTOp_(o: TOp, xs: Prod, ds: Prod, T=None) -> Prod:
    """`gradTOp'` (Types.hs:124)."""
    with saved_activations():
        return o.grad_(_T(xs, T), list(xs), list(ds))


def gradTOp(o: TOp, xs: Prod, T=None) -> Prod:
    """`gradTOp` (Types.hs:127-132): cotangent of the scalar output seeded with 1."""
    T = _T(xs, T)
    with saved_activations():
        return o.grad_(T, list(xs), [T.konst((), 1.0, xs[0])])


# ------------------------------------------------------------------ Category / routing (Types.hs:135-264)
def compose(o2: TOp, o1: TOp) -> TOp:
    """`(.)`: g3 xs ds = g1 xs (g2 (f1 xs) ds)  (Types.hs:141-157).  The reference re-evaluates `f1 xs` inside the gradient;
    here that forward is taken from the saved activations of the enclosing gradient evaluation when it has already run
    (see `saved_activations`), so a chain of n ops costs O(n) launches instead of O(n^2)."""
    assert o1.n_out == o2.n_in, f"cannot compose: {o1.n_out} outputs into {o2.n_in} inputs"
    fused = _fuse(o2, o1)
    if fused is not None:
        return fused
    return TOp(lambda T, xs: o2.run(T, _run_saved(o1, T, xs)),
               lambda T, xs, ds: o1.grad_(T, xs, o2.grad_(T, _run_saved(o1, T, xs), ds)),
               o1.n_in, o2.n_out, ("seq", o1.tag, o2.tag))


def _fuse(o2: TOp, o1: TOp) -> Optional[TOp]:
    """Rewrites applied when two primitives are composed (the deferred evaluator's fusion rules; tags describe structure).
    gmul lM lO lN >>> sumRows, lM >= 1: the row sum commutes with the contraction — a backend that offers `gmulSumRows` evaluates
    the pair (and its VJP) as one primitive; any other backend evaluates the plain composition."""
    if o1.tag[:1] == ("gmul",) and o2.tag == ("sumRows",) and o1.tag[1] >= 1:
        _, lM, lO, lN = o1.tag

        def run(T, xs):
            if hasattr(T, "gmulSumRows"):
                return [T.gmulSumRows(lM, lO, lN, xs[0], xs[1])]
            return o2.run(T, _run_saved(o1, T, xs))

        def grad(T, xs, ds):
            if hasattr(T, "gmulSumRowsVJP"):
                return list(T.gmulSumRowsVJP(lM, lO, lN, xs[0], xs[1], ds[0]))
            return o1.grad_(T, xs, o2.grad_(T, _run_saved(o1, T, xs), ds))
        return TOp(run, grad, 2, 1, ("gmul_sumRows", lM, lO, lN))
    return None


def idOp(n: int) -> TOp:
    return TOp(lambda T, xs: xs, lambda T, xs, ds: ds, n, n, ("id", n))


def firstOp(o: TOp, n_rest: int) -> TOp:
    """`firstOp` (Types.hs:165-181)."""
    a, b = o.n_in, o.n_out
    return TOp(lambda T, xs: o.run(T, xs[:a]) + xs[a:],
               lambda T, xs, ds: o.grad_(T, xs[:a], ds[:b]) + ds[b:],
               a + n_rest, b + n_rest, ("first", o.tag, n_rest))


def secondOp(n_skip: int, o: TOp) -> TOp:
    """`secondOp` (Types.hs:183-199)."""
    return TOp(lambda T, xs: xs[:n_skip] + o.run(T, xs[n_skip:]),
               lambda T, xs, ds: ds[:n_skip] + o.grad_(T, xs[n_skip:], ds[n_skip:]),
               n_skip + o.n_in, n_skip + o.n_out, ("second", n_skip, o.tag))


def then_first(t1: TOp, t2: TOp) -> TOp:
    """`t1 *>> t2 = firstOp t1 >>> t2` (Types.hs:202-209)."""
    return compose(t2, firstOp(t1, t2.n_in - t1.n_out))


def par(o1: TOp, o2: TOp) -> TOp:
    """`(***)` (Types.hs:221-240)."""
    a, c = o1.n_in, o1.n_out
    return TOp(lambda T, xs: o1.run(T, xs[:a]) + o2.run(T, xs[a:]),
               lambda T, xs, ds: o1.grad_(T, xs[:a], ds[:c]) + o2.grad_(T, xs[a:], ds[c:]),
               a + o2.n_in, c + o2.n_out, ("par", o1.tag, o2.tag))


def fanout(o1: TOp, o2: TOp) -> TOp:
    """`(&&&)` (Types.hs:242-264): gradients of the two branches are summed with `sumT`."""
    b = o1.n_out
    return TOp(lambda T, xs: o1.run(T, xs) + o2.run(T, xs),
               lambda T, xs, ds: [T.sumT([g1, g2]) for g1, g2 in _zip(o1.grad_(T, xs, ds[:b]), o2.grad_(T, xs, ds[b:]))],
               o1.n_in, b + o2.n_out, ("fanout", o1.tag, o2.tag))


# ------------------------------------------------------------------ primitives (TOp.hs)
def gradLift(T, f: Callable, f_grad: Optional[Callable], xs: Sequence, dtdy) -> Prod:
    """`TT.gradLift` (Tensor.hs:119-129): for input k, liftT (\\(d:x) -> d * (vfGrad f x)_k) (dtdy : xs).
    `f_grad` is the explicit gradient of a `VFunc` (returns one derivative per input); when absent the derivative is
    taken symbolically, the role `ad` plays in the reference (TOp.hs:212,246)."""
    n = len(xs)
    outs = []
    for k in range(n):
        if f_grad is not None:
            g = lambda d, *x, k=k: d * f_grad(*x)[k]
        else:
            def g(d, *x, k=k):
                # d is variable 0, x_j is variable j+1 of the lifted program
                e = E.Expr.lift(f(*x))
                return d * E.diff(e, k + 1)
        outs.append(T.liftT(g, [dtdy, *xs]))
    return outs


def liftOp(n: int, f: Callable, f_grad: Optional[Callable] = None, name: str = "lift") -> TOp:
    """`liftOp` (TOp.hs:42-54), n >= 1."""
    return TOp(lambda T, xs: [T.liftT(f, xs)],
               lambda T, xs, ds: gradLift(T, f, f_grad, xs, ds[0]), n, 1, ("lift", name, n))


def map_(f: Callable, fprime: Callable, name: str = "map'") -> TOp:
    """`map'` (TOp.hs:198-205)."""
    return liftOp(1, f, lambda x: [fprime(x)], name)


def map(f: Callable, name: str = "map") -> TOp:   # noqa: A001  (name follows the reference)
    """`map f = map' f (diff f)` (TOp.hs:208-213)."""
    return liftOp(1, f, None, name)


def zipN(n: int, f: Callable, name: str = "zipN") -> TOp:
    """`zipN` (TOp.hs:240-247)."""
    return liftOp(n, f, None, name)


def zip(f: Callable, name: str = "zip") -> TOp:   # noqa: A001
    """`zip` (TOp.hs:263-267)."""
    return zipN(2, f, name)


def gmul(lM: int, lO: int, lN: int) -> TOp:
    """`TO.gmul` (TOp.hs:56-94): both VJPs are `gmul`s of the cotangent against a full transpose."""
    def g(T, xs, ds):
        x, y = xs
        dtdz = ds[0]
        return [T.gmul(lM, lN, lO, dtdz, T.transp(y)), T.gmul(lO, lM, lN, T.transp(x), dtdz)]
    return TOp(lambda T, xs: [T.gmul(lM, lO, lN, xs[0], xs[1])], g, 2, 1, ("gmul", lM, lO, lN))


def inner(lM: int, lN: int) -> TOp: return gmul(lM, 1, lN)     # TOp.hs:304-311
def outer(lM: int, lN: int) -> TOp: return gmul(lM, 0, lN)     # TOp.hs:313-320
def dot() -> TOp: return inner(0, 0)                            # TOp.hs:322-325
def matVec() -> TOp: return inner(1, 0)                         # TOp.hs:327-331
def vecMat() -> TOp: return inner(0, 1)                         # TOp.hs:333-337
def matMat() -> TOp: return inner(1, 1)                         # TOp.hs:339-343


def transpOp() -> TOp:
    """`transpOp` (TOp.hs:97-103)."""
    return TOp(lambda T, xs: [T.transp(xs[0])], lambda T, xs, ds: [T.transp(ds[0])], 1, 1, ("transp",))


def sumRows() -> TOp:
    """`sumRows` (TOp.hs:151-159): VJP = mapRows (LS LZ) (const dtdz)."""
    return TOp(lambda T, xs: [T.sumRows(xs[0])],
               lambda T, xs, ds: [T.broadcastRows(xs[0].shape[0], ds[0])], 1, 1, ("sumRows",))


def sumOp(n: int) -> TOp:
    """`sumOp` (TOp.hs:161-169)."""
    return TOp(lambda T, xs: [T.sumT(xs)], lambda T, xs, ds: [ds[0]] * len(xs), n, 1, ("sumOp", n))


def scale(alpha: float) -> TOp:
    """`scale` (TOp.hs:171-176)."""
    return TOp(lambda T, xs: [T.scaleT(alpha, xs[0])], lambda T, xs, ds: [T.scaleT(alpha, ds[0])], 1, 1, ("scale", alpha))


def negate() -> TOp:
    """`negate = scale (-1)` (TOp.hs:194-195)."""
    return scale(-1.0)


def konst(shapes: Sequence[Sequence[int]], v: float) -> TOp:
    """`konst` (TOp.hs:185-192): needs the backend passed to runTOp (no inputs to infer it from)."""
    return TOp(lambda T, xs: [T.konst(tuple(s), v, None) for s in shapes], lambda T, xs, ds: [], 0, len(shapes), ("konst", v))


def add() -> TOp:
    """`add` (TOp.hs:215-220)."""
    return TOp(lambda T, xs: [T.sumT(xs)], lambda T, xs, ds: [ds[0], ds[0]], 2, 1, ("add",))


def add3() -> TOp:
    """`add3` (TOp.hs:222-229)."""
    return TOp(lambda T, xs: [T.sumT(xs)], lambda T, xs, ds: [ds[0]] * 3, 3, 1, ("add3",))


def replicate(n: int) -> TOp:
    """`replicate` (TOp.hs:287-293)."""
    return TOp(lambda T, xs: [xs[0]] * n, lambda T, xs, ds: [T.sumT(ds)], 1, n, ("replicate", n))


def duplicate() -> TOp:
    """`duplicate` (TOp.hs:295-301)."""
    return TOp(lambda T, xs: [xs[0], xs[0]], lambda T, xs, ds: [T.sumT([ds[0], ds[1]])], 1, 2, ("duplicate",))


def swap() -> TOp:
    """`swap` (TOp.hs:346-351)."""
    return TOp(lambda T, xs: [xs[1], xs[0]], lambda T, xs, ds: [ds[1], ds[0]], 2, 2, ("swap",))


def shuffle(idx: Sequence[int], n_in: int) -> TOp:
    """`shuffle` (TOp.hs:106-134): output j = input idx[j]; the gradient of input i is the sum of the cotangents of
    every output that selected it (an unselected input gets a zero tensor)."""
    def g(T, xs, ds):
        out = []
        for i in range(n_in):
            picks = [d for j, d in _zip(idx, ds) if j == i]
            out.append(T.sumT(picks) if picks else T.scaleT(0.0, xs[i]))
        return out
    return TOp(lambda T, xs: [xs[j] for j in idx], g, n_in, len(idx), ("shuffle", tuple(idx)))


def swap_(nN: int, nM: int) -> TOp:
    """`swap'` (TOp.hs:353-357): (ns ++ ms) -> (ms ++ ns)."""
    return shuffle(list(range(nN, nN + nM)) + list(range(nN)), nN + nM)


def drop(n: int, n_in: int) -> TOp:
    """`drop` (TOp.hs:359-369)."""
    return shuffle(list(range(n, n_in)), n_in)


def take(n: int, n_in: int) -> TOp:
    """`take` (TOp.hs:371-381)."""
    return shuffle(list(range(n)), n_in)
