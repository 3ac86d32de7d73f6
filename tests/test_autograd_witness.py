"""A witness for the oracle that its author did not derive: torch.autograd in fp64 on the CPU.

The reference has no fixtures (test/Spec.hs:1-2), so every VJP in oracle/tensor_ops_oracle.py comes from one reading of
TOp.hs:56-94 / NeuralNet.hs:38-77.  Here only the FORWARD functions are written down (with torch primitives, from the
definitions in the reference's source: Types.hs:60-73 for gmul/transp, NeuralNet.hs:42-77 for logistic/softmax/losses,
FeedForward.hs:209-212 for ffLayer) and reverse-mode autograd produces the gradients — an independent derivation the oracle's
hand-written VJPs and the TOp chain rule (`g1 xs (g2 (f1 xs) ds)`, Types.hs:135-157) must agree with to ~1e-12.
"""
import itertools
import string

import numpy as np
import pytest
import torch

from oracle import tensor_ops_oracle as O

torch.set_default_dtype(torch.float64)


def T(a, grad=True):
    return torch.tensor(np.asarray(a, dtype=np.float64), requires_grad=grad)


def agree(got, want, what, tol=1e-11):
    got = np.asarray(got, dtype=np.float64); want = np.asarray(want.detach().numpy() if hasattr(want, "detach") else want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    err = np.linalg.norm(got - want) / max(np.linalg.norm(want), 1e-300)
    assert err < tol, f"{what}: oracle vs autograd rel err {err:.3e}"


# ------------------------------------------------------------------ ffLayer' >>> logistic with a supplied cotangent (config 2 semantics)
@pytest.mark.parametrize("seed,B,i,o", [(0, 7, 5, 3), (1, 33, 17, 9), (2, 64, 48, 40)])
def test_fflayer_logistic_vjp(seed, B, i, o):
    rng = np.random.default_rng(seed)
    X = rng.uniform(-1, 1, (B, i)); W = rng.normal(0, 0.5, (o, i)); b = rng.normal(0, 0.5, o); dA = rng.normal(size=(B, o))
    tX, tW, tb = T(X), T(W), T(b)
    A = 1.0 / (1.0 + torch.exp(-(tX @ tW.T + tb)))          # FeedForward.hs:209-212, NeuralNet.hs:42-44
    (A * T(dA, False)).sum().backward()                      # VJP with cotangent dA; parameter gradients summed over samples
    for got in (O.fflayer_logistic_dense(X, W, b, dA), O.fflayer_logistic_per_sample(X, W, b, dA), O.cpu_fflayer_step_reference(X, W, b, dA)):
        agree(got[0], A, "A"); agree(got[1], tX.grad, "dX"); agree(got[2], tW.grad, "dW"); agree(got[3], tb.grad, "db")


# ------------------------------------------------------------------ MLP + loss heads (config 3 semantics)
def torch_mlp_loss(tX, tWs, tbs, acts, loss, tY):
    h = tX
    for W, b, act in zip(tWs, tbs, acts):
        z = h @ W.T + b
        if act == "logistic":
            h = 1.0 / (1.0 + torch.exp(-z))
        elif act == "softmax":                                # NeuralNet.hs:52-59: exp / sum exp, no max-subtraction
            e = torch.exp(z)
            h = e / e.sum(dim=1, keepdim=True)
        else:
            h = z
    if loss == "squaredError":                                # NeuralNet.hs:61-68
        return h, ((tY - h) ** 2).sum()
    return h, -(torch.log(h) * tY).sum()                      # NeuralNet.hs:70-77


@pytest.mark.parametrize("acts,loss", [(["logistic", "logistic", "softmax"], "crossEntropy"), (["logistic", "logistic", "logistic"], "squaredError"),
                                       (["logistic", "softmax"], "squaredError"), (["id", "logistic"], "crossEntropy")])
def test_mlp_netgrad(acts, loss):
    rng = np.random.default_rng(len(acts) * 7 + len(loss))
    dims = [11, 9, 6][: len(acts)] + [5]
    B = 13
    Ws = [rng.normal(0, 0.5, (dims[l + 1], dims[l])) for l in range(len(acts))]
    bs = [rng.normal(0, 0.5, dims[l + 1]) for l in range(len(acts))]
    X = rng.uniform(0, 1, (B, dims[0]))
    Y = np.eye(dims[-1])[rng.integers(0, dims[-1], B)] if loss == "crossEntropy" else rng.uniform(0, 1, (B, dims[-1]))
    tX, tWs, tbs = T(X), [T(w) for w in Ws], [T(b) for b in bs]
    A, L = torch_mlp_loss(tX, tWs, tbs, acts, loss, T(Y, False))
    L.backward()
    oA, oL, odX, odWs, odbs = O.mlp_dense_fwd_grad(X, Ws, bs, acts, loss, Y)
    agree(oA, A, "A"); agree(np.array(oL), L, "loss"); agree(odX, tX.grad, "dX")
    for l in range(len(acts)):
        agree(odWs[l], tWs[l].grad, f"dW{l}"); agree(odbs[l], tbs[l].grad, f"db{l}")


def test_per_sample_topgraph_netgrad_softmax_ce():
    """The TOp machinery itself (genNet's composed TOp + crossEntropy, evaluated by gradTOp with the reference's chain rule) against
    autograd on one sample — not the dense closed form."""
    rng = np.random.default_rng(4)
    net = O.genNet(8, [(6, O.actLogistic), (5, O.actLogistic)], 4, O.softmax, rng)
    x = rng.uniform(0, 1, 8); y = np.eye(4)[2]
    g = O.netGrad(O.crossEntropy(), x, y, net)
    tx = T(x[None, :]); tps = [T(p) for p in net.params]
    _, L = torch_mlp_loss(tx, tps[0::2], tps[1::2], ["logistic", "logistic", "softmax"], "crossEntropy", T(y[None, :], False))
    L.backward()
    agree(g[0], tx.grad[0], "dx")
    for k, p in enumerate(tps):
        agree(g[1 + k], p.grad, f"param {k}")


# ------------------------------------------------------------------ general contraction (Types.hs:60-73): reversed axes on the right operand
def torch_gmul(lM, lO, lN, x, y):
    L = string.ascii_lowercase
    ms, os_, ns = L[:lM], L[lM:lM + lO], L[lM + lO:lM + lO + lN]
    return torch.einsum(f"{ms}{os_},{os_[::-1]}{ns}->{ms}{ns}", x, y)


@pytest.mark.parametrize("lM,lO,lN", [t for t in itertools.product(range(3), range(3), range(3)) if 0 < sum(t) <= 4])
def test_gmul_vjp(lM, lO, lN):
    rng = np.random.default_rng(lM * 9 + lO * 3 + lN)
    ms = tuple(rng.integers(2, 5, size=lM)); os_ = tuple(rng.integers(2, 5, size=lO)); ns = tuple(rng.integers(2, 5, size=lN))
    x = rng.normal(size=ms + os_); y = rng.normal(size=tuple(reversed(os_)) + ns); ct = rng.normal(size=ms + ns)
    tx, ty = T(x), T(y)
    z = torch_gmul(lM, lO, lN, tx, ty)
    (z * T(ct, False)).sum().backward()
    agree(O.gmul(lM, lO, lN, x, y), z, "gmul")
    dx, dy = O.gradTOp_(O.op_gmul(lM, lO, lN), [x, y], [ct])
    agree(dx, tx.grad, "dx"); agree(dy, ty.grad, "dy")


def test_config5_rank3_inner_then_sumrows():
    """BASELINE configs[4] as SURVEY §8-d reads it: inner (LS (LS LZ)) (LS LZ) on x[a,b,c], y[c,n] then sumRows, VJP with a [b,n] cotangent."""
    rng = np.random.default_rng(5)
    x = rng.normal(size=(6, 5, 4)); y = rng.normal(size=(4, 7)); ct = rng.normal(size=(5, 7))
    tx, ty = T(x), T(y)
    out = torch.einsum("abc,cn->abn", tx, ty).sum(dim=0)
    (out * T(ct, False)).sum().backward()
    op = O.op_gmul(2, 1, 1) >> O.op_sumRows()
    agree(O.runTOp(op, [x, y])[0], out, "out")
    dx, dy = O.gradTOp_(op, [x, y], [ct])
    agree(dx, tx.grad, "dx"); agree(dy, ty.grad, "dy")


def test_transp_sumrows_scale_add_routing():
    """transp = full axis reversal (Types.hs:67-73) and the routing combinators' VJPs (fan-out sums cotangents, Types.hs:237-264)."""
    rng = np.random.default_rng(6)
    x = rng.normal(size=(3, 4, 5)); ct = rng.normal(size=(5, 4, 3))
    tx = T(x)
    z = tx.permute(2, 1, 0) * 2.0 + tx.permute(2, 1, 0)
    (z * T(ct, False)).sum().backward()
    op = O.op_transp() >> O.fanout(O.op_scale(2.0), O.idOp(1)) >> O.op_add()
    agree(O.runTOp(op, [x])[0], z, "fwd")
    agree(O.gradTOp_(op, [x], [ct])[0], tx.grad, "dx")
