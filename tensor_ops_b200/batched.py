"""`BatchT`: a second `Tensor` instance that evaluates a per-sample TOp for a whole batch at once.

The reference has no batched code path: `ffLayer` is `TOp '[ '[i], '[o,i], '[o] ] '[ '[o] ]` on single vectors and
training is a fold of per-sample steps (src/TensorOps/Learn/NeuralNet/FeedForward.hs:201-214, app/Dots.hs:74-80).
`BatchT` keeps the reference's closures untouched and changes only the instance they are run at: every value
carries a hidden leading batch axis in HBM (`batched=True`) or is shared by all samples (`batched=False`,
parameters).  Contractions against a shared operand become ONE tensor-core GEMM over the batch.  The per-sample
outer product that a `matVec` VJP produces for the weight gradient (`ger dtdz x`) is never
materialised as [B,o,i]: it stays a lazy pair of factors and is reduced over the batch as `dZ^T X` — a single
split-K GEMM — when the gradient of a shared input is requested (`grad_batched`).

Anything the vectorised rules below do not cover falls back to a per-sample loop of CuTensor calls (still on the
device, counted in `BatchT.fallbacks` so tests can assert the hot path never takes it).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import numpy as np

from . import top as TO
from .tensor import CuTensor


class BatchT:
    fallbacks = 0

    def __init__(self, t: Optional[CuTensor], batched: bool, lazy=None, B: Optional[int] = None):
        self._t, self.batched, self.lazy = t, batched, lazy
        self._B = B

    # ---- structure
    @property
    def t(self) -> CuTensor:
        if self._t is None:
            self._t = self._materialise()
            self.lazy = None
        return self._t

    @property
    def B(self) -> int:
        if self._B is not None:
            return self._B
        return self.t.shape[0] if self.batched else 0

    @property
    def shape(self):
        """per-sample shape"""
        if self.lazy is not None:
            a, c = self.lazy
            return a.shape + c.shape
        s = self._t.shape
        return s[1:] if self.batched else s

    def _materialise(self) -> CuTensor:
        a, c = self.lazy   # per-sample outer product of two batched values
        return _loop(lambda x, y: CuTensor.gmul(len(a.shape), 0, len(c.shape), x, y), [a, c]).t

    # ---- helpers
    @staticmethod
    def _ones(like: CuTensor, n: int) -> CuTensor:
        return like.ctx.full((n,), 1.0)

    @staticmethod
    def _as_batched(x: "BatchT", B: int) -> CuTensor:
        if x.batched:
            return x.t
        return CuTensor.broadcastRows(B, x.t)

    @staticmethod
    def _flat2(t: CuTensor) -> CuTensor:
        s = t.shape
        return t.reshape((s[0], int(np.prod(s[1:])) if len(s) > 1 else 1))

    # ================================================================= class Tensor (Types.hs:52-109)
    @staticmethod
    def liftT(f: Callable, xs: Sequence["BatchT"], like=None) -> "BatchT":
        if not any(x.batched for x in xs):
            return BatchT(CuTensor.liftT(f, [x.t for x in xs]), False)
        B = next(x.B for x in xs if x.batched)
        return BatchT(CuTensor.liftT(f, [BatchT._as_batched(x, B) for x in xs]), True)

    @staticmethod
    def gmul(lM: int, lO: int, lN: int, x: "BatchT", y: "BatchT") -> "BatchT":
        if not x.batched and not y.batched:
            return BatchT(CuTensor.gmul(lM, lO, lN, x.t, y.t), False)
        if x.batched and not y.batched:
            # the batch is one more leading `ms` axis:  z[b,m..,n..] = sum_o x[b,m..,o..] y[rev o.., n..]
            return BatchT(CuTensor.gmul(lM + 1, lO, lN, x.t, y.t), True)
        if not x.batched and y.batched:
            if lN == 0 and lM <= 1 and lO <= 1:
                # z[b,m] = sum_o x[m,o] y[b,o]  ==  Y X^T : (B as ms) x (transp x as [o,m])
                return BatchT(CuTensor.gmul(1, lO, lM, y.t, CuTensor.transp(x.t)), True)
            if lM == 0 and lO == 0:
                # shared scalar times per-sample tensor
                return BatchT(CuTensor.gmul(0, 0, 1 + lN, x.t, y.t), True)
            return _loop(lambda a, c: CuTensor.gmul(lM, lO, lN, a, c), [x, y])
        # both per-sample
        B = x.B
        if lO == 0:
            if lM == 0 or lN == 0:
                # per-sample scalar times per-sample tensor (softmax's `outer LZ (LS LZ)`, NeuralNet.hs:58)
                s, v = (x, y) if lM == 0 else (y, x)
                vt = BatchT._flat2(v.t)
                n = vt.shape[1]
                sv = CuTensor.ger(s.t.reshape((B,)), BatchT._ones(s.t, n))          # [B,n]: s[b] along the row
                out = CuTensor.liftT(lambda p, q: p * q, [sv, vt])
                return BatchT(out.reshape((B,) + tuple(v.shape)), True)
            return BatchT(None, True, lazy=(x, y), B=B)                               # lazy per-sample outer product
        if lM == 0 and lN == 0 and lO <= 1:
            # per-sample dot:  z[b] = sum_o x[b,o] y[b,o]
            prod = CuTensor.liftT(lambda p, q: p * q, [BatchT._flat2(x.t), BatchT._flat2(y.t)])
            return BatchT(CuTensor.gemv(prod, BatchT._ones(prod, prod.shape[1])), True)
        return _loop(lambda a, c: CuTensor.gmul(lM, lO, lN, a, c), [x, y])

    @staticmethod
    def sumT(xs: Sequence["BatchT"]) -> "BatchT":
        if len(xs) == 1:
            return xs[0]
        if not any(x.batched for x in xs):
            return BatchT(CuTensor.sumT([x.t for x in xs]), False)
        B = next(x.B for x in xs if x.batched)
        return BatchT(CuTensor.sumT([BatchT._as_batched(x, B) for x in xs]), True)

    @staticmethod
    def scaleT(a: float, x: "BatchT") -> "BatchT":
        if x.lazy is not None:
            p, q = x.lazy
            return BatchT(None, True, lazy=(BatchT.scaleT(a, p), q), B=x.B)
        return BatchT(CuTensor.scaleT(a, x.t), x.batched)

    @staticmethod
    def transp(x: "BatchT") -> "BatchT":
        if x.lazy is not None:
            p, q = x.lazy
            if len(p.shape) <= 1 and len(q.shape) <= 1:
                return BatchT(None, True, lazy=(q, p), B=x.B)        # (p ⊗ q)^T = q ⊗ p
        if not x.batched:
            return BatchT(CuTensor.transp(x.t), False)
        if len(x.shape) <= 1:
            return x
        return _loop(lambda a: CuTensor.transp(a), [x])

    @staticmethod
    def sumRows(x: "BatchT") -> "BatchT":
        if not x.batched:
            return BatchT(CuTensor.sumRows(x.t), False)
        if len(x.shape) == 1:
            return BatchT(CuTensor.gemv(x.t, BatchT._ones(x.t, x.shape[0])), True)
        return _loop(lambda a: CuTensor.sumRows(a), [x])

    @staticmethod
    def broadcastRows(n: int, row: "BatchT") -> "BatchT":
        if not row.batched:
            return BatchT(CuTensor.broadcastRows(n, row.t), False)
        if len(row.shape) == 0:
            return BatchT(CuTensor.ger(row.t.reshape((row.B,)), BatchT._ones(row.t, n)), True)
        return _loop(lambda a: CuTensor.broadcastRows(n, a), [row])

    @staticmethod
    def mapRows(lN: int, f, x: "BatchT") -> "BatchT":
        return _loop(lambda a: CuTensor.mapRows(lN, f, a), [x]) if x.batched else BatchT(CuTensor.mapRows(lN, f, x.t), False)

    @staticmethod
    def diag(rank: int, v: "BatchT") -> "BatchT":
        return _loop(lambda a: CuTensor.diag(rank, a), [v]) if v.batched else BatchT(CuTensor.diag(rank, v.t), False)

    @staticmethod
    def getDiag(x: "BatchT") -> "BatchT":
        return _loop(lambda a: CuTensor.getDiag(a), [x]) if x.batched else BatchT(CuTensor.getDiag(x.t), False)

    @staticmethod
    def konst(shape, v: float, like: Optional["BatchT"]) -> "BatchT":
        ref = None
        if like is not None:
            ref = like._t if like._t is not None else like.lazy[0].t
        return BatchT(CuTensor.konst(shape, v, ref), False)


def _loop(fn: Callable[..., CuTensor], xs: List[BatchT]) -> BatchT:
    """Per-sample fallback: run `fn` on each sample's views and stack the results (device-side copies)."""
    BatchT.fallbacks += 1
    if BatchT.fallbacks == 1:
        import warnings
        warnings.warn("tensor_ops_b200.batched: an op without a batched kernel is being evaluated one sample at a time "
                      "(BatchT.fallbacks counts these); the result is exact but launch-bound", RuntimeWarning, stacklevel=3)
    B = next(x.B for x in xs if x.batched)
    outs = [fn(*[(x.t.row(b) if x.batched else x.t) for x in xs]) for b in range(B)]
    shape = outs[0].shape
    n = int(np.prod(shape)) if shape else 1
    ctx = outs[0].ctx
    out = ctx.empty((B,) + tuple(shape))
    import ctypes as C
    from . import _lib as L
    for b, o in enumerate(outs):
        v = out.view(b * n, shape)
        slot = L.c_buf(v.b.value)
        ctx.check(L.lib.tops_axpy(ctx.h, 1.0, o.b, None, C.byref(slot)))
    return BatchT(out, True)


def reduce_over_batch(g: BatchT, B: int) -> CuTensor:
    """Sum of per-sample gradients for an input shared by all samples."""
    if g.lazy is not None:
        p, q = g.lazy
        if len(p.shape) == 1 and len(q.shape) == 1:
            # Σ_b p_b ⊗ q_b = P^T Q : one split-K GEMM, both operands MN-major (A is an O(1) transposed view)
            return CuTensor.gmul(1, 1, 1, CuTensor.transp(p.t), q.t)
    if not g.batched:
        return CuTensor.scaleT(float(B), g.t)      # identical contribution from every sample
    return CuTensor.sumRows(g.t)


def grad_batched(op: TO.TOp, xs: List[BatchT], ds: Optional[List[BatchT]] = None) -> List[CuTensor]:
    """`gradTOp` (or `gradTOp'` with cotangents `ds`) of a per-sample TOp over a batch: per-sample gradients for
    batched inputs, batch-summed gradients for shared inputs."""
    B = next(x.B for x in xs if x.batched)
    if ds is None:
        ref = next(x for x in xs if x.batched).t
        ds = [BatchT(ref.ctx.full((B,), 1.0), True)]
    with TO.saved_activations():      # forwards of composed prefixes run once per gradient evaluation (top.compose)
        gs = op.grad_(BatchT, list(xs), list(ds))
    out = []
    for x, g in zip(xs, gs):
        if x.batched:
            out.append(BatchT._as_batched(g, B) if not g.batched else g.t)
        else:
            out.append(reduce_over_batch(g, B))
    return out


def run_batched(op: TO.TOp, xs: List[BatchT]) -> List[BatchT]:
    return op.run(BatchT, list(xs))
