"""CPU oracle for the tensor-ops hot path (TEST INFRASTRUCTURE — never imported by the product).

This module is a NumPy restatement of the algorithm the reference (mstksg/tensor-ops, Haskell)
runs for `runTOp` / `gradTOp` over an `ffLayer` network with its hmatrix (`HMat`) BLAS backend.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs
may import it.  The shipped path (`tensor_ops_b200`) must never route through here.

PARITY UNPINNED.  The reference ships no tests, golden vectors or fixtures (`test/Spec.hs:1-2`
prints "Test suite not yet implemented"), its RNG is unseeded (`app/Dots.hs:130`), and neither GHC
nor hmatrix exists in this environment, so nothing produced by the reference itself can pin this
restatement.  It is pinned instead by (see tests/test_oracle.py):
  * central finite differences of every primitive VJP and of whole networks,
  * three independent restatements of the general contraction agreeing with each other
    (`gmul` = einsum-style tensordot, `gmul_naive` = the element loop of `Data/Nested.hs:451-473`,
    `gmul_btensor` = the BLAS rank-dispatch of `Backend/BTensor.hs:592-716` onto the `HMat` ops),
  * the per-sample reference op sequence agreeing with the dense batched closed form,
  * the behavioural check of `app/Dots.hs` (two-circle target is learned).

All citations are `path:line` into /root/reference.  Tensors are NumPy arrays whose `.shape` is the
type-level dimension list (row-major, first index outermost, as `Data.Nested` nests vectors).
A "Prod" (the tensor stack a TOp consumes/produces) is a Python list of arrays.
"""
from __future__ import annotations

import itertools
import math
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

Prod = List[np.ndarray]

# --------------------------------------------------------------------------------------------
# class Tensor (src/TensorOps/Types.hs:52-109) restated on ndarrays
# --------------------------------------------------------------------------------------------


def liftT(f: Callable[..., np.ndarray], xs: Sequence[np.ndarray]) -> np.ndarray:
    """`liftT` (Types.hs:56-59): apply an n-ary scalar function elementwise to n same-shape tensors."""
    return np.asarray(f(*xs))


def gmul(lM: int, lO: int, lN: int, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """`gmul` (Types.hs:60-66; semantics Data/Nested.hs:451-473).

    x : ms ++ os,  y : Reverse os ++ ns  ->  ms ++ ns
    z[m.., n..] = sum_{o..} x[m.., o..] * y[reverse(o..), n..]
    (the contraction indices of `y` appear in REVERSE order).
    """
    assert x.ndim == lM + lO and y.ndim == lO + lN
    # bring y's leading (reversed) contraction axes into x's order
    perm = list(range(lO - 1, -1, -1)) + list(range(lO, lO + lN))
    yr = np.transpose(y, perm)
    assert x.shape[lM:] == yr.shape[:lO], (x.shape, y.shape, lM, lO, lN)
    return np.tensordot(x, yr, axes=lO)


def gmul_naive(lM: int, lO: int, lN: int, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """Element-at-a-time restatement of `Data.Nested.gmul'` (Nested.hs:451-473): for every `ms` slice of
    x, fold over its `os` indices i, summing x'[i] * (y indexed at reverse(i))."""
    ms, os_ = x.shape[:lM], x.shape[lM:]
    ns = y.shape[lO:]
    z = np.zeros(ms + ns, dtype=np.result_type(x, y))
    for m in itertools.product(*[range(d) for d in ms]):
        acc = np.zeros(ns, dtype=z.dtype)
        for o in itertools.product(*[range(d) for d in os_]):
            acc = acc + x[m + o] * y[tuple(reversed(o))]
        z[m] = acc
    return z


def sumT(xs: Sequence[np.ndarray]) -> np.ndarray:
    """`sumT` (Types.hs:69) = `sum'` = foldl1' (+) (Data/List/Util.hs:7-10) — left fold, in order."""
    acc = xs[0]
    for x in xs[1:]:
        acc = acc + x
    return acc


def scaleT(a, x: np.ndarray) -> np.ndarray:
    """`scaleT` (Types.hs:70)."""
    return (x.dtype.type(a) * x) if x.dtype.kind == "f" else a * x


def transp(x: np.ndarray) -> np.ndarray:
    """`transp` (Types.hs:71-73; Nested.hs:476-528): reverse ALL axes."""
    return np.transpose(x)


def mapRows(lN: int, f: Callable[[np.ndarray], np.ndarray], x: np.ndarray) -> np.ndarray:
    """`mapRows` (Types.hs:77-81): apply f to every sub-tensor under the leading lN axes."""
    out = np.empty_like(x)
    for i in itertools.product(*[range(d) for d in x.shape[:lN]]):
        out[i] = f(x[i])
    return out


def sumRows(x: np.ndarray) -> np.ndarray:
    """`sumRows` (Types.hs:82-84): sum over the leading axis."""
    return x.sum(axis=0)


def diag(rank: int, v: np.ndarray) -> np.ndarray:
    """`diag` (Types.hs:85-88): vector [n] -> rank-`rank` tensor [n,..,n] with v on the generalised diagonal."""
    n = v.shape[0]
    out = np.zeros((n,) * rank, dtype=v.dtype)
    idx = np.arange(n)
    out[(idx,) * rank] = v
    return out


def getDiag(x: np.ndarray) -> np.ndarray:
    """`getDiag` (Types.hs:89-92)."""
    n = x.shape[0]
    idx = np.arange(n)
    return x[(idx,) * x.ndim].copy()


def konst(shape: Sequence[int], v, dtype=np.float64) -> np.ndarray:
    """`TT.konst` (Tensor.hs:49-54) = generate (const v)."""
    return np.full(tuple(shape), v, dtype=dtype)


# --------------------------------------------------------------------------------------------
# instance BLAS (HMat a) (src/TensorOps/BLAS/HMat.hs:103-231): op-for-op, pass-for-pass
# --------------------------------------------------------------------------------------------


def hm_axpy(alpha, x, y=None):
    """HMat.hs:135-139: `maybe id (add y) . scale alpha $ x` — scale pass then add pass."""
    r = x.dtype.type(alpha) * x
    return r if y is None else y + r


def hm_dot(x, y):
    """HMat.hs:141-142."""
    return np.dot(x, y)


def hm_ger(x, y):
    """HMat.hs:144-145: `x outer y`."""
    return np.outer(x, y)


def hm_gemv(alpha, a, x, beta_y=None):
    """HMat.hs:147-153: `(a #>) . scale alpha $ x`, then `add (scale beta y)`."""
    r = a @ (x.dtype.type(alpha) * x)
    if beta_y is not None:
        beta, y = beta_y
        r = (y.dtype.type(beta) * y) + r
    return r


def hm_gemm(alpha, a, b, beta_c=None):
    """HMat.hs:154-160: `(a <>) . scale alpha $ b`, then `add (scale beta c)`."""
    r = a @ (b.dtype.type(alpha) * b)
    if beta_c is not None:
        beta, c = beta_c
        r = (c.dtype.type(beta) * c) + r
    return r


def hm_scale(alpha, x):
    """HMat.hs:161."""
    return x.dtype.type(alpha) * x


def hm_eye(n, dtype):
    """HMat.hs:215."""
    return np.eye(n, dtype=dtype)


def hm_trace(x):
    """HMat.hs:221-222: sumElements . takeDiag."""
    return np.diagonal(x).sum()


def bt_add(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """`(+)` of `Num (BTensor v b ns)` (BTensor.hs:110-113): scalars add; vectors `axpy 1 x (Just y)`;
    matrices `gemm 1 x eye (Just (1,y))` (an O(n m^2) GEMM — reproduced faithfully for rounding order);
    rank>=3 recurse over the leading axis (zipBase, BTensor.hs:476-496)."""
    if x.ndim == 0:
        return x + y
    if x.ndim == 1:
        return hm_axpy(1, x, y)
    if x.ndim == 2:
        return hm_gemm(1, x, hm_eye(x.shape[1], x.dtype), (1, y))
    return np.stack([bt_add(a, b) for a, b in zip(x, y)])


def gmul_btensor(lM: int, lO: int, lN: int, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """`BTensor.gmulB`/`gmulBLAS`/`dispatchBLAS`/`naiveGMul` (BTensor.hs:141-175,592-716): the rank-case
    analysis that turns one `gmul` into hmatrix BLAS calls."""
    if lN >= 2 or lO >= 3:                                   # :615-616
        return gmul_naive(lM, lO, lN, x, y)
    if lO == 2:
        if lN == 0:                                          # :611-613 trace(gemm xs ys) per trailing matrix
            lead = x.shape[:lM]
            out = np.empty(lead, dtype=x.dtype)
            for m in itertools.product(*[range(d) for d in lead]):
                out[m] = hm_trace(hm_gemm(1, x[m], y))
            return out
        return gmul_naive(lM, lO, lN, x, y)                  # :614
    if lO == 1:
        def one(xs):                                         # dispatchBLAS :151-175
            if xs.ndim == 1 and lN == 0:
                return np.asarray(hm_dot(xs, y))
            if xs.ndim == 1 and lN == 1:
                return hm_gemv(1, np.transpose(y), xs)       # vector-matrix :160-163
            if xs.ndim == 2 and lN == 0:
                return hm_gemv(1, xs, y)                     # matrix-vector :169-171
            return hm_gemm(1, xs, y)                         # matrix-matrix :172-174
        if lM <= 1:                                          # :694
            return one(x)
        lead = x.shape[: lM - 1]                             # mapBTM :695-713
        res = [one(x[m]) for m in itertools.product(*[range(d) for d in lead])]
        return np.stack(res).reshape(lead + res[0].shape)
    # lO == 0
    if lM == 0:
        return hm_axpy(x, y) if lN == 1 else np.asarray(x * y)   # scalar-vector / scalar-scalar :146-150
    if lM == 1:
        return hm_axpy(y, x) if lN == 0 else hm_ger(x, y)        # vector-scalar / ger :164-168
    if lN == 0:                                              # :674-675,683-687 scaleB on trailing matrices
        return hm_scale(y, x)
    return gmul_naive(lM, lO, lN, x, y)                      # :680,692


# --------------------------------------------------------------------------------------------
# forward-mode dual numbers (stands in for `ad`'s `diff`/`grad`, TOp.hs:212,246)
# --------------------------------------------------------------------------------------------


class Dual:
    __array_priority__ = 1000

    def __init__(self, v, d):
        self.v, self.d = v, d

    @staticmethod
    def lift(x):
        return x if isinstance(x, Dual) else Dual(x, 0.0)

    def __add__(self, o): o = Dual.lift(o); return Dual(self.v + o.v, self.d + o.d)
    __radd__ = __add__
    def __sub__(self, o): o = Dual.lift(o); return Dual(self.v - o.v, self.d - o.d)
    def __rsub__(self, o): return Dual.lift(o) - self
    def __mul__(self, o): o = Dual.lift(o); return Dual(self.v * o.v, self.d * o.v + self.v * o.d)
    __rmul__ = __mul__
    def __truediv__(self, o): o = Dual.lift(o); return Dual(self.v / o.v, (self.d * o.v - self.v * o.d) / (o.v * o.v))
    def __rtruediv__(self, o): return Dual.lift(o) / self
    def __neg__(self): return Dual(-self.v, -self.d)
    def exp(self): e = np.exp(self.v); return Dual(e, self.d * e)
    def log(self): return Dual(np.log(self.v), self.d / self.v)
    def sqrt(self): s = np.sqrt(self.v); return Dual(s, self.d / (2 * s))
    def tanh(self): t = np.tanh(self.v); return Dual(t, self.d * (1 - t * t))


def diff(f):
    """`diff f` from `ad` (used by `TO.map`, TOp.hs:208-213)."""
    def df(x):
        r = f(Dual(x, np.ones_like(x)))
        return np.broadcast_to(np.asarray(Dual.lift(r).d, dtype=x.dtype), x.shape).copy()
    return df


def grad_n(f, n):
    """`grad f` from `ad` for an n-ary scalar function (used by `TO.zipN`, TOp.hs:240-247)."""
    def gf(*xs):
        outs = []
        for k in range(n):
            args = [Dual(x, np.ones_like(x) if j == k else np.zeros_like(x)) for j, x in enumerate(xs)]
            outs.append(np.broadcast_to(np.asarray(Dual.lift(f(*args)).d, dtype=xs[0].dtype), xs[0].shape).copy())
        return outs
    return gf


# --------------------------------------------------------------------------------------------
# data TOp, Category, routing combinators (src/TensorOps/Types.hs:122-264)
# --------------------------------------------------------------------------------------------


@dataclass
class TOp:
    """`data TOp ns ms` (Types.hs:122-125): a forward closure and a VJP closure. `n_in`/`n_out` are the
    lengths of the type-level tensor-stack lists (needed where the reference uses `Known Length`)."""
    run: Callable[[Prod], Prod]
    grad_: Callable[[Prod, Prod], Prod]
    n_in: int
    n_out: int

    def __rshift__(self, other: "TOp") -> "TOp":      # f >>> g
        return compose(other, self)


def runTOp(o: TOp, xs: Prod) -> Prod:
    return o.run(list(xs))


def gradTOp_(o: TOp, xs: Prod, ds: Prod) -> Prod:
    """`gradTOp'` (Types.hs:124)."""
    return o.grad_(list(xs), list(ds))


def gradTOp(o: TOp, xs: Prod) -> Prod:
    """`gradTOp` (Types.hs:127-132): seed the scalar output's cotangent with 1."""
    return o.grad_(list(xs), [np.ones((), dtype=xs[0].dtype)])


def compose(o2: TOp, o1: TOp) -> TOp:
    """`(.)` of `Category TOp` (Types.hs:141-157): g3 xs ds = g1 xs (g2 (f1 xs) ds) — the forward of the
    first op is RECOMPUTED inside the gradient."""
    assert o1.n_out == o2.n_in, (o1.n_out, o2.n_in)
    return TOp(lambda xs: o2.run(o1.run(xs)),
               lambda xs, ds: o1.grad_(xs, o2.grad_(o1.run(xs), ds)),
               o1.n_in, o2.n_out)


def idOp(n: int) -> TOp:
    """`id` (Types.hs:136-138)."""
    return TOp(lambda xs: xs, lambda xs, ds: ds, n, n)


def firstOp(o: TOp, n_rest: int) -> TOp:
    """`firstOp` (Types.hs:165-181): apply `o` to the first n_in entries, pass `n_rest` through."""
    a, b = o.n_in, o.n_out
    return TOp(lambda xs: o.run(xs[:a]) + xs[a:],
               lambda xs, ds: o.grad_(xs[:a], ds[:b]) + ds[b:],
               a + n_rest, b + n_rest)


def secondOp(n_skip: int, o: TOp) -> TOp:
    """`secondOp` (Types.hs:183-199)."""
    return TOp(lambda xs: xs[:n_skip] + o.run(xs[n_skip:]),
               lambda xs, ds: ds[:n_skip] + o.grad_(xs[n_skip:], ds[n_skip:]),
               n_skip + o.n_in, n_skip + o.n_out)


def then_first(t1: TOp, t2: TOp) -> TOp:
    """`t1 *>> t2 = firstOp t1 >>> t2` (Types.hs:202-209)."""
    return compose(t2, firstOp(t1, t2.n_in - t1.n_out))


def par(o1: TOp, o2: TOp) -> TOp:
    """`(***)` (Types.hs:221-240)."""
    a, c = o1.n_in, o1.n_out
    return TOp(lambda xs: o1.run(xs[:a]) + o2.run(xs[a:]),
               lambda xs, ds: o1.grad_(xs[:a], ds[:c]) + o2.grad_(xs[a:], ds[c:]),
               a + o2.n_in, c + o2.n_out)


def fanout(o1: TOp, o2: TOp) -> TOp:
    """`(&&&)` (Types.hs:242-264): VJP = sumT [g1, g2] per input."""
    b = o1.n_out
    return TOp(lambda xs: o1.run(xs) + o2.run(xs),
               lambda xs, ds: [sumT([g1, g2]) for g1, g2 in zip(o1.grad_(xs, ds[:b]), o2.grad_(xs, ds[b:]))],
               o1.n_in, b + o2.n_out)


# --------------------------------------------------------------------------------------------
# primitive TOps (src/TensorOps/TOp.hs)
# --------------------------------------------------------------------------------------------


def gradLift(f_grad: Callable[..., List[np.ndarray]], xs: Sequence[np.ndarray], dtdy: np.ndarray):
    """`TT.gradLift` (Tensor.hs:119-129): for input k, liftT (\\(d:x) -> d * (vfGrad f x)_k) (dtdy:xs);
    vfGrad is re-evaluated once per input, as the reference does."""
    return [liftT(lambda d, *x, k=k: d * f_grad(*x)[k], [dtdy, *xs]) for k in range(len(xs))]


def op_liftOp(n: int, f, f_grad) -> TOp:
    """`liftOp` (TOp.hs:42-54) for n >= 1."""
    return TOp(lambda xs: [liftT(f, xs)],
               lambda xs, ds: gradLift(f_grad, xs, ds[0]), n, 1)


def op_map_(f, fprime) -> TOp:
    """`map'` (TOp.hs:198-205)."""
    return op_liftOp(1, f, lambda x: [fprime(x)])


def op_map(f) -> TOp:
    """`map f = map' f (diff f)` (TOp.hs:208-213)."""
    return op_map_(f, diff(f))


def op_zipN(n: int, f) -> TOp:
    """`zipN u f = zipN' u f (grad f)` (TOp.hs:240-247)."""
    return op_liftOp(n, f, grad_n(f, n))


def op_gmul(lM: int, lO: int, lN: int, gm=gmul) -> TOp:
    """`TO.gmul` (TOp.hs:56-94): VJPs are themselves `gmul`s of the cotangent with full transposes."""
    def g(xs, ds):
        x, y = xs
        dtdz = ds[0]
        dx = gm(lM, lN, lO, dtdz, transp(y))
        dy = gm(lO, lM, lN, transp(x), dtdz)
        return [dx, dy]
    return TOp(lambda xs: [gm(lM, lO, lN, xs[0], xs[1])], g, 2, 1)


def op_inner(lM, lN, gm=gmul): return op_gmul(lM, 1, lN, gm)      # TOp.hs:304-311
def op_outer(lM, lN, gm=gmul): return op_gmul(lM, 0, lN, gm)      # TOp.hs:313-320
def op_dot(gm=gmul): return op_inner(0, 0, gm)                    # TOp.hs:322-325
def op_matVec(gm=gmul): return op_inner(1, 0, gm)                 # TOp.hs:327-331
def op_vecMat(gm=gmul): return op_inner(0, 1, gm)                 # TOp.hs:333-337
def op_matMat(gm=gmul): return op_inner(1, 1, gm)                 # TOp.hs:339-343


def op_transp() -> TOp:
    """`transpOp` (TOp.hs:97-103)."""
    return TOp(lambda xs: [transp(xs[0])], lambda xs, ds: [transp(ds[0])], 1, 1)


def op_sumRows() -> TOp:
    """`sumRows` (TOp.hs:151-159): VJP broadcasts dtdz over the leading axis via mapRows (const dtdz)."""
    return TOp(lambda xs: [sumRows(xs[0])],
               lambda xs, ds: [mapRows(1, lambda _r: ds[0], xs[0])], 1, 1)


def op_sumOp(n: int) -> TOp:
    """`sumOp` (TOp.hs:161-169)."""
    return TOp(lambda xs: [sumT(xs)], lambda xs, ds: [ds[0] for _ in xs], n, 1)


def op_scale(alpha, add=None) -> TOp:
    """`scale` (TOp.hs:171-176)."""
    return TOp(lambda xs: [scaleT(alpha, xs[0])], lambda xs, ds: [scaleT(alpha, ds[0])], 1, 1)


def op_negate() -> TOp:
    """`negate = scale (-1)` (TOp.hs:194-195)."""
    return op_scale(-1)


def op_konst(shapes: Sequence[Sequence[int]], v, dtype=np.float64) -> TOp:
    """`konst` (TOp.hs:185-192)."""
    return TOp(lambda xs: [konst(s, v, dtype) for s in shapes], lambda xs, ds: [], 0, len(shapes))


def op_add(add=None) -> TOp:
    """`add` (TOp.hs:215-220): forward sumT [x,y]; VJP (dtdz, dtdz)."""
    plus = (lambda xs: sumT(xs)) if add is None else (lambda xs: add(xs[0], xs[1]))
    return TOp(lambda xs: [plus(xs)], lambda xs, ds: [ds[0], ds[0]], 2, 1)


def op_replicate(n: int) -> TOp:
    """`replicate` (TOp.hs:287-293)."""
    return TOp(lambda xs: [xs[0]] * n, lambda xs, ds: [sumT(ds)], 1, n)


def op_duplicate(add=None) -> TOp:
    """`duplicate` (TOp.hs:295-301): VJP sumT [d1,d2]."""
    plus = (lambda a, b: sumT([a, b])) if add is None else add
    return TOp(lambda xs: [xs[0], xs[0]], lambda xs, ds: [plus(ds[0], ds[1])], 1, 2)


def op_swap() -> TOp:
    """`swap` (TOp.hs:346-351)."""
    return TOp(lambda xs: [xs[1], xs[0]], lambda xs, ds: [ds[1], ds[0]], 2, 2)


# --------------------------------------------------------------------------------------------
# activations and losses (src/TensorOps/Learn/NeuralNet.hs)
# --------------------------------------------------------------------------------------------


def logistic(x):
    """NeuralNet.hs:42-44: 1 / (1 + exp (-x))."""
    return 1 / (1 + np.exp(-x)) if not isinstance(x, Dual) else 1 / (1 + (-x).exp())


def logistic_(x):
    """NeuralNet.hs:46-50: logix * (1 - logix) with logix RECOMPUTED from x."""
    s = logistic(x)
    return s * (1 - s)


def _exp(x): return x.exp() if isinstance(x, Dual) else np.exp(x)
def _log(x): return x.log() if isinstance(x, Dual) else np.log(x)
def _recip(x): return 1 / x


def actLogistic() -> TOp:
    """NeuralNet.hs:38-40."""
    return op_map_(logistic, logistic_)


def softmax(gm=gmul, add=None) -> TOp:
    """NeuralNet.hs:52-59: map exp >>> duplicate >>> firstOp (sumRows >>> map recip) >>> outer LZ (LS LZ).
    No max-subtraction (numerically naive on purpose: it is what the reference computes)."""
    return (op_map(_exp) >> op_duplicate(add)
            >> firstOp(op_sumRows() >> op_map(_recip), 1) >> op_outer(0, 1, gm))


def squaredError(gm=gmul, add=None) -> TOp:
    """NeuralNet.hs:61-68: negate *>> add >>> duplicate >>> dot   on (a, target)."""
    return then_first(op_negate(), op_add(add) >> op_duplicate(add) >> op_dot(gm))


def crossEntropy(gm=gmul) -> TOp:
    """NeuralNet.hs:71-77: map log *>> dot >>> negate   on (a, target)."""
    return then_first(op_map(_log), op_dot(gm) >> op_negate())


# --------------------------------------------------------------------------------------------
# feed-forward networks (src/TensorOps/Learn/NeuralNet/FeedForward.hs)
# --------------------------------------------------------------------------------------------


@dataclass
class Network:
    """`Network t i o = N sing TOp params` (FeedForward.hs:57-61)."""
    op: TOp                 # TOp ('[i] : ps) '[ '[o] ]
    params: Prod


def ffLayer_(gm=gmul, add=None) -> TOp:
    """`ffLayer'` (FeedForward.hs:209-212): firstOp (swap >>> matVec) >>> add   on (x, W, b)."""
    return firstOp(op_swap() >> op_matVec(gm), 1) >> op_add(add)


def ffLayer(i: int, o: int, rng: np.random.Generator, dtype=np.float64, gm=gmul, add=None) -> Network:
    """`ffLayer` (FeedForward.hs:201-214): W[o,i], b[o] ~ N(0, 0.5^2), W drawn first, row-major."""
    w = rng.normal(0.0, 0.5, size=(o, i)).astype(dtype)
    b = rng.normal(0.0, 0.5, size=(o,)).astype(dtype)
    return Network(ffLayer_(gm, add), [w, b])


def net_then_act(n: Network, f: TOp) -> Network:
    """`(*~)` (FeedForward.hs:103-108)."""
    return Network(n.op >> f, n.params)


def net_compose(n1: Network, n2: Network) -> Network:
    """`(~*~)` (FeedForward.hs:82-90): o1 *>> o2 with params appended."""
    return Network(then_first(n1.op, n2.op), n1.params + n2.params)


def genNet(i: int, hidden: Sequence[Tuple[int, Callable[[], TOp]]], o: int, out_act: Callable[[], TOp],
           rng: np.random.Generator, dtype=np.float64, gm=gmul, add=None) -> Network:
    """`genNet` (FeedForward.hs:216-235). The recursion `go` generates the TAIL network first and the
    layer in front of it afterwards (`n <- go sl xs; l <- ffLayer g`), so random draws happen
    last-layer-first; reproduced here."""
    def go(j, xs):
        if not xs:
            return net_then_act(ffLayer(j, o, rng, dtype, gm, add), out_act())
        (h, act), rest = xs[0], xs[1:]
        n = go(h, rest)
        l = ffLayer(j, h, rng, dtype, gm, add)
        return net_compose(net_then_act(l, act()), n)
    return go(i, list(hidden))


def runNetwork(n: Network, x: np.ndarray) -> np.ndarray:
    """FeedForward.hs:123-129."""
    return runTOp(n.op, [x] + n.params)[0]


def netGrad(loss: TOp, x: np.ndarray, y: np.ndarray, n: Network) -> Prod:
    """`netGrad` (FeedForward.hs:178-199): gradTOp (o *>> loss) (x :< p >: y), target's gradient dropped.
    Returns [dx, dp...]."""
    o_ = then_first(n.op, loss)
    return gradTOp(o_, [x] + n.params + [y])[:-1]


def trainNetwork(loss: TOp, r, x, y, n: Network) -> Network:
    """`trainNetwork` (FeedForward.hs:131-148): p' = zip (\\o g -> o - r*g) p grad."""
    g = netGrad(loss, x, y, n)[1:]
    rr = n.params[0].dtype.type(r)
    return Network(n.op, [p - rr * gp for p, gp in zip(n.params, g)])


# --------------------------------------------------------------------------------------------
# auto-encoders (src/TensorOps/Learn/NeuralNet/AutoEncoder.hs) — SURVEY §8-f4
# --------------------------------------------------------------------------------------------


@dataclass
class Encoder:
    """`Encoder t i o = E { eEncoder :: Network t i o, eDecoder :: Network t o i }` (AutoEncoder.hs:36-39)."""
    enc: Network
    dec: Network


def encoderNet(e: Encoder) -> Network:
    """`encoderNet (E e d) = e >>> d` (AutoEncoder.hs:80-84; `>>>` on Networks is `~*~`)."""
    return net_compose(e.enc, e.dec)


def encode(e: Encoder, x): return runNetwork(e.enc, x)             # AutoEncoder.hs:41-47
def decode(e: Encoder, h): return runNetwork(e.dec, h)             # AutoEncoder.hs:49-55
def encodeDecode(e: Encoder, x): return runNetwork(encoderNet(e), x)   # AutoEncoder.hs:57-62


def _encoder_loss_op(loss: TOp, net: Network) -> TOp:
    """firstOp duplicate >>> secondOp @'[ '[i] ] o >>> swap >>> loss   (AutoEncoder.hs:72-78, 129-138): the input is both
    the network's input and the loss target."""
    return firstOp(op_duplicate(), len(net.params)) >> secondOp(1, net.op) >> op_swap() >> loss


def testEncoder(loss: TOp, e: Encoder, x) -> float:
    """AutoEncoder.hs:64-78."""
    net = encoderNet(e)
    return float(runTOp(_encoder_loss_op(loss, net), [x] + net.params)[0])


def encGrad(loss: TOp, x, e: Encoder):
    """`encGrad` (AutoEncoder.hs:110-142): gradients of the reconstruction loss w.r.t. (encoder params, decoder params)."""
    net = encoderNet(e)
    gr = gradTOp(_encoder_loss_op(loss, net), [x] + net.params)[1:]
    nE = len(e.enc.params)
    return gr[:nE], gr[nE:]


def trainEncoder(loss: TOp, r, x, e: Encoder) -> Encoder:
    """`trainEncoder` (AutoEncoder.hs:86-108): p' = p - r*g on both halves."""
    gE, gD = encGrad(loss, x, e)
    rr = e.enc.params[0].dtype.type(r)
    return Encoder(Network(e.enc.op, [p - rr * g for p, g in zip(e.enc.params, gE)]),
                   Network(e.dec.op, [p - rr * g for p, g in zip(e.dec.params, gD)]))


# --------------------------------------------------------------------------------------------
# recurrent networks (src/TensorOps/Learn/NeuralNet/Recurrent.hs) — SURVEY §8-f4
# --------------------------------------------------------------------------------------------


def op_shuffle(idx: Sequence[int], n_in: int) -> TOp:
    """`shuffle` (TOp.hs:106-134): output j = input idx[j]; the cotangent of input i is the sum of the cotangents of every
    output that selected it (zeros if none did)."""
    def g(xs, ds):
        out = []
        for i in range(n_in):
            picks = [d for j, d in zip(idx, ds) if j == i]
            out.append(sumT(picks) if picks else np.zeros_like(xs[i]))
        return out
    return TOp(lambda xs: [xs[j] for j in idx], g, n_in, len(idx))


def op_swap_(nN: int, nM: int) -> TOp:
    """`swap'` (TOp.hs:353-357): (ns ++ ms) -> (ms ++ ns)."""
    return op_shuffle(list(range(nN, nN + nM)) + list(range(nN)), nN + nM)


def op_drop(n: int, n_in: int) -> TOp:
    """`drop` (TOp.hs:359-369)."""
    return op_shuffle(list(range(n, n_in)), n_in)


def op_take(n: int, n_in: int) -> TOp:
    """`take` (TOp.hs:371-381)."""
    return op_shuffle(list(range(n)), n_in)


def op_add3() -> TOp:
    """`add3` (TOp.hs:222-229)."""
    return TOp(lambda xs: [sumT(xs)], lambda xs, ds: [ds[0]] * 3, 3, 1)


@dataclass
class RNetwork:
    """`Network t i o = N { _nOp :: TOp ('[i] : ss ++ ps) ('[o] : ss), _nState :: Prod t ss, _nParams :: Prod t ps }`
    (Recurrent.hs:69-75)."""
    op: TOp
    state: Prod
    params: Prod


def r_fullyConnected(i: int, o: int, act: Callable[[], TOp], rng: np.random.Generator, dtype=np.float64) -> RNetwork:
    """`fullyConnected` (Recurrent.hs:97-125): y = W x + W' h + b is the OUTPUT, act(y) the new state.  Draw order s, w, w', b;
    parameter order (w', w, b)."""
    s = rng.normal(0.0, 0.5, size=(o,)).astype(dtype)
    w = rng.normal(0.0, 0.5, size=(o, i)).astype(dtype)
    w_ = rng.normal(0.0, 0.5, size=(o, o)).astype(dtype)
    b = rng.normal(0.0, 0.5, size=(o,)).astype(dtype)
    fc = (secondOp(1, firstOp(op_swap() >> op_matVec(), 2) >> firstOp(op_swap(), 1))
          >> firstOp(op_swap() >> op_matVec(), 2)
          >> op_add3()
          >> op_duplicate()
          >> secondOp(1, act()))
    return RNetwork(fc, [s], [w_, w, b])


def r_stateless(n: Network) -> RNetwork:
    """`stateless` (Recurrent.hs:127-137)."""
    return RNetwork(n.op, [], list(n.params))


def r_then_act(n: RNetwork, f: TOp) -> RNetwork:
    """`(*~)` (Recurrent.hs:262-267): o >>> firstOp f."""
    return RNetwork(n.op >> firstOp(f, len(n.state)), n.state, n.params)


def r_compose(n1: RNetwork, n2: RNetwork) -> RNetwork:
    """`(~*~)` (Recurrent.hs:178-233): states ss2 ++ ss1, parameters ps1 ++ ps2."""
    s1, p1, s2, p2 = len(n1.state), len(n1.params), len(n2.state), len(n2.params)
    o = (secondOp(1, firstOp(op_swap_(s2, s1 + p1), p2))
         >> firstOp(n1.op, s2 + p2)
         >> secondOp(1, op_swap_(s1, s2 + p2))
         >> firstOp(n2.op, s1))
    return RNetwork(o, n2.state + n1.state, n1.params + n2.params)


def r_runNetwork(n: RNetwork, x):
    """`runNetwork` (Recurrent.hs:235-244): returns (output, network with the new state)."""
    out = runTOp(n.op, [x] + n.state + n.params)
    return out[0], RNetwork(n.op, out[1:], n.params)


def r_unroll(nS: int, nP: int, o: TOp, n: int) -> TOp:
    """`unroll` (Recurrent.hs:392-431): TOp (Replicate n '[i] ++ ss ++ ps) (ss ++ Replicate n '[o]).  The step is applied to the
    LAST input of the list first, and its output is appended LAST — so with inputs in reverse time order (netGrad passes
    `reverse xs`) the outputs come out in reverse time order too."""
    if n == 0:
        return op_take(nS, nS + nP)
    m = n - 1
    step = fanout(o, op_drop(1 + nS, 1 + nS + nP)) >> op_swap_(1, nS + nP)      # (x, ss, ps) -> (ss', ps, y)
    return secondOp(m, step) >> firstOp(r_unroll(nS, nP, o, m), 1)


def r_rollup(loss: TOp, n: int) -> TOp:
    """`rollup` (Recurrent.hs:434-463): TOp (Replicate n '[o] ++ Replicate n '[o]) '[ '[] ] — the LAST output is paired with the
    FIRST target, and the per-step losses are added."""
    if n == 0:
        return op_konst([()], 0.0)
    if n == 1:
        return loss
    m = n - 1
    return secondOp(m, firstOp(loss, m) >> op_swap_(1, m)) >> firstOp(r_rollup(loss, m), 1) >> op_add()


def r_netGrad(loss: TOp, xs: Sequence[np.ndarray], ys: Sequence[np.ndarray], n: RNetwork):
    """`netGrad` (Recurrent.hs:277-324): back-propagation through time by unrolling.  Returns (input gradients, state gradients,
    parameter gradients); as in the reference the input gradients are in the order of `reverse xs` (the reference converts the
    gradient Prod of the reversed inputs back to a Vec without re-reversing it)."""
    T, nS, nP = len(xs), len(n.state), len(n.params)
    unrolled = r_unroll(nS, nP, n.op, T) >> op_drop(nS, nS + T)
    full = firstOp(unrolled, T) >> r_rollup(loss, T)
    grad = gradTOp(full, list(xs)[::-1] + n.state + n.params + list(ys))[:T + nS + nP]
    return grad[:T], grad[T:T + nS], grad[T + nS:]


def r_trainNetwork(loss: TOp, rS, rP, xs, ys, n: RNetwork) -> RNetwork:
    """`trainNetwork'` (Recurrent.hs:326-352): separate rates for the initial state and the parameters."""
    _, gS, gP = r_netGrad(loss, xs, ys, n)
    return RNetwork(n.op, [s - s.dtype.type(rS) * g for s, g in zip(n.state, gS)], [p - p.dtype.type(rP) * g for p, g in zip(n.params, gP)])


# --------------------------------------------------------------------------------------------
# batched semantics fixed by SURVEY §8(d): per-sample runTOp + gradTOp', parameter grads summed
# --------------------------------------------------------------------------------------------


def fflayer_logistic_per_sample(X, W, b, dA, gm=gmul, add=None):
    """Reference-faithful evaluation: for each sample, runTOp and gradTOp' of `ffLayer' >>> logistic`
    with cotangent dA[s]; parameter gradients accumulated over samples in order."""
    op = ffLayer_(gm, add) >> actLogistic()
    A = np.empty((X.shape[0], W.shape[0]), dtype=X.dtype)
    dX = np.empty_like(X)
    dW = np.zeros_like(W)
    db = np.zeros_like(b)
    for s in range(X.shape[0]):
        A[s] = runTOp(op, [X[s], W, b])[0]
        gx, gw, gb = gradTOp_(op, [X[s], W, b], [dA[s]])
        dX[s] = gx
        dW += gw
        db += gb
    return A, dX, dW, db


def fflayer_logistic_dense(X, W, b, dA):
    """Dense closed form of the same thing: Z = X W^T + b, A = σ(Z), dZ = dA ⊙ A(1-A),
    dW = dZ^T X, db = Σ_s dZ, dX = dZ W."""
    Z = X @ W.T + b
    A = logistic(Z)
    dZ = dA * (A * (1 - A))
    return A, dZ @ W, dZ.T @ X, dZ.sum(axis=0)


def mlp_dense_fwd_grad(X, Ws, bs, acts, loss, Y):
    """Dense batched netGrad for a genNet-style MLP (acts: 'logistic'|'softmax'|'id' per layer;
    loss: 'squaredError'|'crossEntropy'), per-sample losses summed.  Returns (A_out, loss_sum, dX, dWs, dbs).
    Derived from the per-sample TOp VJPs above and checked against them in tests/test_oracle.py."""
    hs = [X]
    outs = []
    for W, b, act in zip(Ws, bs, acts):
        Z = hs[-1] @ W.T + b
        if act == "logistic":
            A = logistic(Z)
        elif act == "softmax":
            E = np.exp(Z)
            A = E * (1 / E.sum(axis=1, keepdims=True))
        else:
            A = Z
        outs.append((Z, A))
        hs.append(A)
    A = hs[-1]
    if loss == "squaredError":
        D = Y - A
        L = (D * D).sum()
        dA = -2 * D
    else:
        L = -(np.log(A) * Y).sum()
        dA = -(Y / A)
    dWs, dbs = [], []
    for li in range(len(Ws) - 1, -1, -1):
        Z, Aout = outs[li]
        act = acts[li]
        if act == "logistic":
            dZ = dA * (Aout * (1 - Aout))
        elif act == "softmax":
            # VJP of exp -> (sum -> recip) -> scalar*vector, NeuralNet.hs:52-59
            E = np.exp(Z)
            r = 1 / E.sum(axis=1, keepdims=True)
            dE = dA * r + (-(r * r)) * (dA * E).sum(axis=1, keepdims=True)
            dZ = dE * E
        else:
            dZ = dA
        dWs.append(dZ.T @ hs[li])
        dbs.append(dZ.sum(axis=0))
        dA = dZ @ Ws[li]
    return A, L, dA, dWs[::-1], dbs[::-1]


# --------------------------------------------------------------------------------------------
# app/Dots.hs behaviour (config 1)
# --------------------------------------------------------------------------------------------


def dots_target(v: np.ndarray) -> float:
    """Dots.hs:65-69,93-100: inside either circle of radius 0.33 centred at (0.33,0.33) / (-0.33,-0.33)."""
    def inc(c):
        d = v - c
        return float(d @ d) <= 0.33 ** 2
    return 1.0 if inc(np.full(2, 0.33)) or inc(np.full(2, -0.33)) else 0.0


def dots_train(n_samples=50000, hidden=(16,), rate=1.0, seed=0, dtype=np.float64, gm=gmul):
    """`netTest` (Dots.hs:60-82): U(-1,1)^2 inputs, two-circle target, per-sample SGD fold."""
    rng = np.random.default_rng(seed)
    inps = rng.uniform(-1, 1, size=(n_samples, 2)).astype(dtype)
    outs = np.array([[dots_target(v)] for v in inps], dtype=dtype)
    net = genNet(2, [(h, actLogistic) for h in hidden], 1, actLogistic, rng, dtype, gm)
    loss = squaredError(gm)
    for x, y in zip(inps, outs):
        net = trainNetwork(loss, rate, x, y, net)
    return net


def dots_accuracy(net: Network, n=2000, seed=123) -> float:
    rng = np.random.default_rng(seed)
    pts = rng.uniform(-1, 1, size=(n, 2)).astype(net.params[0].dtype)
    ok = 0
    for v in pts:
        ok += (float(runNetwork(net, v)[0]) > 0.5) == (dots_target(v) > 0.5)
    return ok / n


# --------------------------------------------------------------------------------------------
# CPU baseline kernels for bench.py (the reference's op sequence, timed; not a target)
# --------------------------------------------------------------------------------------------


def cpu_fflayer_step_reference(X, W, b, dA):
    """One fwd+grad over a batch exactly as the hmatrix backend would execute it per sample
    (SURVEY §3.2-3.3): runTOp = gemv, axpy, cmap logistic; gradTOp' recomputes the forward
    (Types.hs:155), then elementwise d*σ'(z), ger for dW, gemv (tr W) for dx."""
    Wt = np.ascontiguousarray(W)           # tr W is an O(1) view in hmatrix (HMat.hs:175)
    dW = np.zeros_like(W)
    db = np.zeros_like(b)
    A = np.empty((X.shape[0], W.shape[0]), dtype=X.dtype)
    dX = np.empty_like(X)
    one = X.dtype.type(1)
    for s in range(X.shape[0]):
        x = X[s]
        z = hm_axpy(one, hm_gemv(one, W, x), b)            # runTOp
        A[s] = logistic(z)
        z2 = hm_axpy(one, hm_gemv(one, W, x), b)           # recomputed forward inside gradTOp'
        dz = dA[s] * logistic_(z2)                         # gradLift
        db += dz
        dW += hm_ger(dz, x)                                # dispatchOut
        dX[s] = hm_gemv(one, Wt.T, dz)                     # dispatchMV on transp W
    return A, dX, dW, db
