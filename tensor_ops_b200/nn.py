"""Feed-forward networks over TOps: the mirror of src/TensorOps/Learn/NeuralNet.hs and
src/TensorOps/Learn/NeuralNet/FeedForward.hs, plus the batched, fused device entry points.

Per-sample functions (`runNetwork`, `netGrad`, `trainNetwork`) have the reference's exact semantics and run every
tensor method on the device.  The `*Batched` functions evaluate the same TOp for a whole batch with parameter
gradients summed over samples (the batched semantics fixed in SURVEY §8-d); when the network is a recognised
ffLayer chain they dispatch to the fused kernels (`tops_mlp_fwd_grad`), otherwise to the generic `BatchT` instance.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

from . import _lib as L
from . import expr as E
from . import top as TO
from .tensor import Context, CuTensor, default_context

# ------------------------------------------------------------------ activations / losses (NeuralNet.hs)


@dataclass
class Activation:
    """`newtype Activation` (NeuralNet.hs:15-19) + the kernel-catalogue id the fused path uses."""
    op: Callable[[], TO.TOp]
    kind: int   # L.ACT_* or -1 when only the generic path can run it


def actMap(f: Callable) -> Activation:
    """NeuralNet.hs:21-25."""
    return Activation(lambda: TO.map(f), -1)


def actMap_(f: Callable, fprime: Callable) -> Activation:
    """NeuralNet.hs:27-32."""
    return Activation(lambda: TO.map_(f, fprime), -1)


logistic = E.logistic      # NeuralNet.hs:42-44
logistic_ = E.logistic_    # NeuralNet.hs:46-50

actLogistic = Activation(lambda: TO.map_(logistic, logistic_, "logistic"), L.ACT_LOGISTIC)   # NeuralNet.hs:38-40
actId = Activation(lambda: TO.idOp(1), L.ACT_ID)


def softmax() -> TO.TOp:
    """NeuralNet.hs:52-59: map exp >>> duplicate >>> firstOp (sumRows >>> map recip) >>> outer LZ (LS LZ)."""
    op = (TO.map(E.exp, "exp") >> TO.duplicate() >> TO.firstOp(TO.sumRows() >> TO.map(E.recip, "recip"), 1) >> TO.outer(0, 1))
    op.tag = ("softmax",)
    return op


actSoftmax = Activation(softmax, L.ACT_SOFTMAX)   # NeuralNet.hs:34-36


def squaredError() -> TO.TOp:
    """NeuralNet.hs:61-68: negate *>> add >>> duplicate >>> dot."""
    op = TO.then_first(TO.negate(), TO.add() >> TO.duplicate() >> TO.dot())
    op.tag = ("squaredError",)
    return op


def crossEntropy() -> TO.TOp:
    """NeuralNet.hs:71-77: map log *>> dot >>> negate."""
    op = TO.then_first(TO.map(E.log, "log"), TO.dot() >> TO.negate())
    op.tag = ("crossEntropy",)
    return op


_LOSS_KIND = {("squaredError",): L.LOSS_SQUARED_ERROR, ("crossEntropy",): L.LOSS_CROSS_ENTROPY}

# ------------------------------------------------------------------ networks (FeedForward.hs)


@dataclass
class Network:
    """`Network t i o = N sing TOp params` (FeedForward.hs:57-61).  `layers` records, for networks assembled by
    `ffLayer` / `genNet`, the activation kind of each dense layer so the batched path can fuse them; it is None for
    networks built from arbitrary TOps."""
    op: TO.TOp
    params: List[CuTensor]
    layers: Optional[List[int]] = None


def ffLayer_() -> TO.TOp:
    """`ffLayer'` (FeedForward.hs:209-212): firstOp (swap >>> matVec) >>> add   on (x, W, b)."""
    op = TO.firstOp(TO.swap() >> TO.matVec(), 1) >> TO.add()
    op.tag = ("ffLayer",)
    return op


def ffLayer(i: int, o: int, seed: int, ctx: Optional[Context] = None) -> Network:
    """`ffLayer` (FeedForward.hs:201-214): W[o,i], b[o] ~ N(0, 0.5^2) drawn on the device (Philox)."""
    ctx = ctx or default_context()
    w = ctx.rand_normal((o, i), 0.0, 0.5, seed * 2 + 1)
    b = ctx.rand_normal((o,), 0.0, 0.5, seed * 2 + 2)
    return Network(ffLayer_(), [w, b], [L.ACT_ID])


def net_then_act(n: Network, act: Activation) -> Network:
    """`(*~)` (FeedForward.hs:103-108)."""
    layers = None
    if n.layers is not None and n.layers[-1] == L.ACT_ID and act.kind >= 0:
        layers = n.layers[:-1] + [act.kind]
    return Network(n.op >> act.op(), n.params, layers)


def net_compose(n1: Network, n2: Network) -> Network:
    """`(~*~)` (FeedForward.hs:82-90)."""
    layers = n1.layers + n2.layers if (n1.layers is not None and n2.layers is not None) else None
    return Network(TO.then_first(n1.op, n2.op), n1.params + n2.params, layers)


def genNet(i: int, hidden: Sequence[Tuple[int, Activation]], o: int, out_act: Activation, seed: int = 0,
           ctx: Optional[Context] = None) -> Network:
    """`genNet` (FeedForward.hs:216-235)."""
    counter = [seed * 1000]

    def go(j, xs):
        if not xs:
            counter[0] += 1
            return net_then_act(ffLayer(j, o, counter[0], ctx), out_act)
        (h, act), rest = xs[0], xs[1:]
        n = go(h, rest)
        counter[0] += 1
        return net_compose(net_then_act(ffLayer(j, h, counter[0], ctx), act), n)
    return go(i, list(hidden))


def networkFromParams(params: Sequence[CuTensor], acts: Sequence[Activation]) -> Network:
    """Assemble a genNet-shaped network from explicit parameters [W0, b0, W1, b1, ...]."""
    net = None
    for l, act in enumerate(acts):
        layer = net_then_act(Network(ffLayer_(), [params[2 * l], params[2 * l + 1]], [L.ACT_ID]), act)
        net = layer if net is None else net_compose(net, layer)
    return net


def runNetwork(n: Network, x: CuTensor) -> CuTensor:
    """FeedForward.hs:123-129."""
    return TO.runTOp(n.op, [x] + n.params)[0]


def netGrad(loss: TO.TOp, x: CuTensor, y: CuTensor, n: Network) -> List[CuTensor]:
    """`netGrad` (FeedForward.hs:178-199): gradTOp (o *>> loss) (x :< p >: y); returns [dx, dparams...]."""
    return TO.gradTOp(TO.then_first(n.op, loss), [x] + n.params + [y])[:-1]


def trainNetwork(loss: TO.TOp, r: float, x: CuTensor, y: CuTensor, n: Network) -> Network:
    """`trainNetwork` (FeedForward.hs:131-148): p' = zip (\\o g -> o - r*g) p grad, via the fused SGD kernel."""
    g = netGrad(loss, x, y, n)[1:]
    return Network(n.op, sgd_step(n.params, g, r), n.layers)


# ------------------------------------------------------------------ fused batched entry points
def _arr(xs: Sequence[CuTensor]):
    return (L.c_buf * max(1, len(xs)))(*[t.b for t in xs])


def sgd_step(params: Sequence[CuTensor], grads: Sequence[CuTensor], rate: float) -> List[CuTensor]:
    ctx = params[0].ctx
    outs = (L.c_buf * len(params))()
    ctx.check(L.lib.tops_sgd_step(ctx.h, len(params), _arr(params), _arr(grads), float(rate), outs))
    return [CuTensor(ctx, L.c_buf(outs[j])) for j in range(len(params))]


def fflayer_fwd(X: CuTensor, W: CuTensor, b: CuTensor, act: int = L.ACT_LOGISTIC) -> CuTensor:
    """runTOp of `ffLayer' >>> act` for every row of X (tops_fflayer_fwd)."""
    out = L.c_buf()
    X.ctx.check(L.lib.tops_fflayer_fwd(X.ctx.h, X.b, W.b, b.b, act, C.byref(out)))
    return CuTensor(X.ctx, out)


def fflayer_fwd_grad(X: CuTensor, W: CuTensor, b: CuTensor, dA: CuTensor, act: int = L.ACT_LOGISTIC, want_dx: bool = True,
                     out: Optional[Sequence[Optional[CuTensor]]] = None):
    """runTOp + gradTOp' of `ffLayer' >>> act` over a batch: returns (A, dX, dW, db); dW/db summed over samples.
    `out` may hold pre-allocated (A, dX, dW, db) tensors to write into (e.g. dW/db views of one packed buffer)."""
    ctx = X.ctx
    slots = [L.c_buf(), L.c_buf(), L.c_buf(), L.c_buf()]
    if out is not None:
        for j, t in enumerate(out):
            if t is not None:
                slots[j] = L.c_buf(t.b.value)
    ctx.check(L.lib.tops_fflayer_fwd_grad(ctx.h, X.b, W.b, b.b, act, dA.b, C.byref(slots[0]),
                                          C.byref(slots[1]) if want_dx else None, C.byref(slots[2]), C.byref(slots[3])))
    res = []
    for j in range(4):
        if out is not None and out[j] is not None:
            res.append(out[j])
        elif j == 1 and not want_dx:
            res.append(None)
        else:
            res.append(CuTensor(ctx, slots[j]))
    return tuple(res)


def fflayer_fwd_grad_mc(X: CuTensor, W: CuTensor, b: CuTensor, dA: CuTensor, grads_multicast_ptr: int, act: int = L.ACT_LOGISTIC,
                        want_dx: bool = True, out: Optional[Sequence[Optional[CuTensor]]] = None):
    """Data-parallel runTOp + gradTOp' with the gradient all-reduce fused into the dW GEMM (tops_fflayer_fwd_grad_mc): finished
    dW regions and db leave the device as NVLS multimem reductions into the packed buffer behind `grads_multicast_ptr` (see
    dp.FusedGradAllReduce for the begin/end protocol).  Returns (A, dX, grads_local); `out` may hold pre-allocated tensors for
    them (grads_local: this rank's own packed [dW‖db] sum, o*i + o floats)."""
    ctx = X.ctx
    slots = [L.c_buf(), L.c_buf(), L.c_buf()]
    if out is not None:
        for j, t in enumerate(out):
            if t is not None:
                slots[j] = L.c_buf(t.b.value)
    ctx.check(L.lib.tops_fflayer_fwd_grad_mc(ctx.h, X.b, W.b, b.b, act, dA.b, C.byref(slots[0]), C.byref(slots[1]) if want_dx else None,
                                             C.byref(slots[2]), C.c_void_p(grads_multicast_ptr)))
    res = []
    for j in range(3):
        if out is not None and j < len(out) and out[j] is not None:
            res.append(out[j])
        elif j == 1 and not want_dx:
            res.append(None)
        else:
            res.append(CuTensor(ctx, slots[j]))
    return tuple(res)


def fflayer_fwd_grad_host(ctx: Context, X_host, W: CuTensor, b: CuTensor, dA_host, act: int = L.ACT_LOGISTIC, grads_out=None,
                          allreduce: Optional[Callable[[], None]] = None, workspace=None, n_chunks: int = 0):
    """Host-buffer entry point of the batched fwd+grad (tops_fflayer_fwd_grad_host): `fromList` the batch (X, dA: C-contiguous
    fp32 host arrays, pinned for full PCIe rate), run the forward + VJP on the device with the host->device copies of row chunk
    k+1 overlapping the GEMMs of chunk k, and `toList` the packed parameter gradient [dW‖db] into `grads_out`.
    `workspace` = (A_dev, dX_dev, packed_dev) re-uses device output buffers between steps; `allreduce` (data-parallel runs) is
    called on the packed device gradient before it is read back."""
    import numpy as np
    B, i = X_host.shape
    o = W.shape[0]
    if X_host.dtype != np.float32 or dA_host.dtype != np.float32 or not X_host.flags.c_contiguous or not dA_host.flags.c_contiguous:
        raise ValueError("fflayer_fwd_grad_host: X and dA must be C-contiguous float32 host arrays")
    if dA_host.shape != (B, o):
        raise ValueError("fflayer_fwd_grad_host: dA must be [B, o]")
    if grads_out is None:
        grads_out = np.empty(o * i + o, dtype=np.float32)
    slots = [L.c_buf(), L.c_buf(), L.c_buf()]
    if workspace is not None:
        for j, t in enumerate(workspace):
            slots[j] = L.c_buf(t.b.value)
    direct = allreduce is None
    ctx.check(L.lib.tops_fflayer_fwd_grad_host(ctx.h, X_host.ctypes.data, dA_host.ctypes.data, B, W.b, b.b, act, n_chunks,
                                               C.byref(slots[0]), C.byref(slots[1]), C.byref(slots[2]),
                                               grads_out.ctypes.data if direct else None))
    outs = list(workspace) if workspace is not None else [CuTensor(ctx, s_) for s_ in slots]
    if not direct:
        allreduce()
        outs[2].download_into(grads_out)
    return grads_out


def fflayer_grad(X: CuTensor, W: CuTensor, b: CuTensor, dA: CuTensor, act: int = L.ACT_LOGISTIC, A_saved: Optional[CuTensor] = None):
    """gradTOp' of the layer (tops_fflayer_grad); recomputes the forward unless A_saved is given."""
    ctx = X.ctx
    dX, dW, db = L.c_buf(), L.c_buf(), L.c_buf()
    ctx.check(L.lib.tops_fflayer_grad(ctx.h, X.b, W.b, b.b, act, dA.b, A_saved.b if A_saved is not None else None,
                                      C.byref(dX), C.byref(dW), C.byref(db)))
    return CuTensor(ctx, dX), CuTensor(ctx, dW), CuTensor(ctx, db)


def mlp_fwd(Ws: Sequence[CuTensor], bs: Sequence[CuTensor], acts: Sequence[int], X: CuTensor) -> CuTensor:
    ctx = X.ctx
    out = L.c_buf()
    a = (C.c_int * len(acts))(*acts)
    ctx.check(L.lib.tops_mlp_fwd(ctx.h, len(Ws), _arr(Ws), _arr(bs), a, X.b, C.byref(out)))
    return CuTensor(ctx, out)


def mlp_fwd_grad(Ws: Sequence[CuTensor], bs: Sequence[CuTensor], acts: Sequence[int], loss: int, X: CuTensor, Y: CuTensor,
                 want_dx: bool = True):
    """Batched netGrad of an ffLayer chain (tops_mlp_fwd_grad): returns (A_out, loss_sum, dX, dWs, dbs)."""
    ctx = X.ctx
    n = len(Ws)
    A, ls, dX = L.c_buf(), L.c_buf(), L.c_buf()
    dW = (L.c_buf * n)()
    db = (L.c_buf * n)()
    a = (C.c_int * n)(*acts)
    ctx.check(L.lib.tops_mlp_fwd_grad(ctx.h, n, _arr(Ws), _arr(bs), a, loss, X.b, Y.b, C.byref(A), C.byref(ls),
                                      C.byref(dX) if want_dx else None, dW, db))
    return (CuTensor(ctx, A), CuTensor(ctx, ls), CuTensor(ctx, dX) if want_dx else None,
            [CuTensor(ctx, L.c_buf(dW[l])) for l in range(n)], [CuTensor(ctx, L.c_buf(db[l])) for l in range(n)])


def runNetworkBatched(n: Network, X: CuTensor) -> CuTensor:
    """`runNetwork` for every row of X.  Fused for ffLayer chains, generic `BatchT` evaluation otherwise."""
    if n.layers is not None:
        return mlp_fwd(n.params[0::2], n.params[1::2], n.layers, X)
    from .batched import BatchT
    return TO.runTOp(n.op, [BatchT(X, True)] + [BatchT(p, False) for p in n.params], BatchT)[0].t


def netGradBatched(loss: TO.TOp, X: CuTensor, Y: CuTensor, n: Network, want_dx: bool = True):
    """`netGrad` over a batch, parameter gradients summed over samples: returns (loss_sum, dX, [dparams...])."""
    kind = _LOSS_KIND.get(loss.tag)
    if n.layers is not None and kind is not None:
        _, ls, dX, dWs, dbs = mlp_fwd_grad(n.params[0::2], n.params[1::2], n.layers, kind, X, Y, want_dx)
        grads = [g for pair in zip(dWs, dbs) for g in pair]
        return ls, dX, grads
    from .batched import BatchT, grad_batched
    full = TO.then_first(n.op, loss)
    xs = [BatchT(X, True)] + [BatchT(p, False) for p in n.params] + [BatchT(Y, True)]
    ls = TO.runTOp(full, xs, BatchT)[0]
    gs = grad_batched(full, xs)
    return CuTensor.sumRows(ls.t) if ls.batched else ls.t, gs[0], gs[1:-1]


def trainNetworkBatched(loss: TO.TOp, r: float, X: CuTensor, Y: CuTensor, n: Network) -> Network:
    """One mini-batch step p' = p - r * Σ_s grad_s (the per-sample rule of FeedForward.hs:141-147 applied to the
    batch-summed gradient)."""
    _, _, grads = netGradBatched(loss, X, Y, n, want_dx=False)
    return Network(n.op, sgd_step(n.params, grads, r), n.layers)
