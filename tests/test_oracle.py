"""Pins the CPU oracle (oracle/tensor_ops_oracle.py).  The reference has no golden vectors
(test/Spec.hs:1-2 is a stub), so the oracle is pinned by finite differences, by three independent
restatements of `gmul` agreeing, by per-sample-vs-dense agreement and by the Dots behaviour check."""
import itertools

import numpy as np
import pytest

from oracle import tensor_ops_oracle as O


def _fd_grad(f, xs, k, eps=1e-6):
    """central finite differences of scalar f(xs) w.r.t. xs[k]"""
    g = np.zeros_like(xs[k])
    it = np.nditer(xs[k], flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        xp = [x.copy() for x in xs]; xp[k][i] += eps
        xm = [x.copy() for x in xs]; xm[k][i] -= eps
        g[i] = (f(xp) - f(xm)) / (2 * eps)
    return g


@pytest.mark.parametrize("lM,lO,lN", [t for t in itertools.product(range(3), range(3), range(3)) if sum(t) <= 4])
def test_gmul_three_witnesses(lM, lO, lN):
    rng = np.random.default_rng(lM * 9 + lO * 3 + lN)
    ms = tuple(rng.integers(2, 4, size=lM)); os_ = tuple(rng.integers(2, 4, size=lO)); ns = tuple(rng.integers(2, 4, size=lN))
    x = rng.normal(size=ms + os_)
    y = rng.normal(size=tuple(reversed(os_)) + ns)
    z = O.gmul(lM, lO, lN, x, y)
    assert z.shape == ms + ns
    np.testing.assert_allclose(z, O.gmul_naive(lM, lO, lN, x, y), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(z, O.gmul_btensor(lM, lO, lN, x, y), rtol=1e-12, atol=1e-12)


def test_gmul_reversed_contraction_order():
    # z[m,n] = sum_{o1,o2} x[m,o1,o2] y[o2,o1,n]   (Types.hs:60-66)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 3, 4)); y = rng.normal(size=(4, 3, 5))
    np.testing.assert_allclose(O.gmul(1, 2, 1, x, y), np.einsum("mab,ban->mn", x, y), rtol=1e-12)


def test_transp_is_full_reversal():
    x = np.arange(24.0).reshape(2, 3, 4)
    assert O.transp(x).shape == (4, 3, 2)
    assert O.transp(x)[3, 1, 0] == x[0, 1, 3]


@pytest.mark.parametrize("lM,lO,lN", [(1, 1, 0), (0, 1, 1), (1, 1, 1), (1, 0, 1), (0, 1, 0), (2, 1, 1), (1, 2, 1), (2, 2, 0)])
def test_gmul_vjp_finite_differences(lM, lO, lN):
    rng = np.random.default_rng(7)
    ms = tuple(rng.integers(2, 4, size=lM)); os_ = tuple(rng.integers(2, 4, size=lO)); ns = tuple(rng.integers(2, 4, size=lN))
    x = rng.normal(size=ms + os_); y = rng.normal(size=tuple(reversed(os_)) + ns)
    ct = rng.normal(size=ms + ns)
    op = O.op_gmul(lM, lO, lN)
    dx, dy = O.gradTOp_(op, [x, y], [ct])
    f = lambda xs: float((O.runTOp(op, xs)[0] * ct).sum())
    np.testing.assert_allclose(dx, _fd_grad(f, [x, y], 0), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(dy, _fd_grad(f, [x, y], 1), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("loss_name", ["squaredError", "crossEntropy"])
def test_network_gradient_finite_differences(loss_name):
    rng = np.random.default_rng(3)
    out_act = O.actLogistic if loss_name == "squaredError" else O.softmax
    net = O.genNet(5, [(4, O.actLogistic), (3, O.actLogistic)], 3, out_act, rng)
    loss = O.squaredError() if loss_name == "squaredError" else O.crossEntropy()
    x = rng.uniform(-1, 1, size=5)
    y = rng.uniform(0, 1, size=3) if loss_name == "squaredError" else np.eye(3)[1]
    g = O.netGrad(loss, x, y, net)
    full = O.then_first(net.op, loss)
    f = lambda xs: float(O.runTOp(full, xs)[0])
    inp = [x] + net.params + [y]
    for k in range(len(g)):
        np.testing.assert_allclose(g[k], _fd_grad(f, inp, k), rtol=2e-6, atol=1e-8)


def test_blas_dispatch_path_matches_plain_path():
    """The BTensor->HMat op sequence (gmul_btensor, bt_add) and the plain restatement give the same network gradient."""
    rng1, rng2 = np.random.default_rng(11), np.random.default_rng(11)
    n1 = O.genNet(6, [(5, O.actLogistic)], 2, O.actLogistic, rng1)
    n2 = O.genNet(6, [(5, O.actLogistic)], 2, O.actLogistic, rng2, gm=O.gmul_btensor, add=O.bt_add)
    x = np.linspace(-1, 1, 6); y = np.array([0.2, 0.9])
    g1 = O.netGrad(O.squaredError(), x, y, n1)
    g2 = O.netGrad(O.squaredError(O.gmul_btensor, O.bt_add), x, y, n2)
    for a, b in zip(g1, g2):
        np.testing.assert_allclose(a, b, rtol=1e-12, atol=1e-14)


def test_per_sample_matches_dense_fflayer():
    rng = np.random.default_rng(5)
    B, i, o = 17, 9, 7
    X = rng.uniform(-1, 1, size=(B, i)); W = rng.normal(0, 0.5, size=(o, i)); b = rng.normal(0, 0.5, size=o)
    dA = rng.normal(size=(B, o))
    ref = O.fflayer_logistic_per_sample(X, W, b, dA)
    ref2 = O.fflayer_logistic_per_sample(X, W, b, dA, gm=O.gmul_btensor, add=O.bt_add)
    ref3 = O.cpu_fflayer_step_reference(X, W, b, dA)
    dense = O.fflayer_logistic_dense(X, W, b, dA)
    for r in (ref, ref2, ref3):
        for a, d in zip(r, dense):
            np.testing.assert_allclose(a, d, rtol=1e-11, atol=1e-13)


@pytest.mark.parametrize("out_act,loss", [("logistic", "squaredError"), ("softmax", "crossEntropy")])
def test_mlp_dense_matches_per_sample_netgrad(out_act, loss):
    rng = np.random.default_rng(9)
    acts = {"logistic": O.actLogistic, "softmax": O.softmax}
    net = O.genNet(8, [(6, O.actLogistic), (5, O.actLogistic)], 4, acts[out_act], rng)
    lossop = O.squaredError() if loss == "squaredError" else O.crossEntropy()
    B = 11
    X = rng.uniform(0, 1, size=(B, 8))
    Y = rng.uniform(0, 1, size=(B, 4)) if loss == "squaredError" else np.eye(4)[rng.integers(0, 4, size=B)]
    Ws, bs = net.params[0::2], net.params[1::2]
    A, L, dX, dWs, dbs = O.mlp_dense_fwd_grad(X, Ws, bs, ["logistic", "logistic", out_act], loss, Y)
    acc = [np.zeros_like(p) for p in net.params]
    Lref = 0.0
    for s in range(B):
        g = O.netGrad(lossop, X[s], Y[s], net)
        np.testing.assert_allclose(g[0], dX[s], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(O.runNetwork(net, X[s]), A[s], rtol=1e-12)
        Lref += float(O.runTOp(O.then_first(net.op, lossop), [X[s]] + net.params + [Y[s]])[0])
        for a, gp in zip(acc, g[1:]):
            a += gp
    np.testing.assert_allclose(L, Lref, rtol=1e-12)
    for li in range(3):
        np.testing.assert_allclose(acc[2 * li], dWs[li], rtol=1e-10, atol=1e-13)
        np.testing.assert_allclose(acc[2 * li + 1], dbs[li], rtol=1e-10, atol=1e-13)


def test_map_uses_ad_style_derivative():
    x = np.linspace(0.5, 2, 7)
    op = O.op_map(lambda v: 1 / (1 + O._exp(-v)))
    (g,) = O.gradTOp_(op, [x], [np.ones_like(x)])
    np.testing.assert_allclose(g, O.logistic_(x), rtol=1e-12)


def test_fanout_and_par_routing():
    x = np.array([1.0, 2.0, 3.0])
    op = O.fanout(O.op_scale(2.0), O.op_scale(3.0))
    assert [a.tolist() for a in O.runTOp(op, [x])] == [[2, 4, 6], [3, 6, 9]]
    (g,) = O.gradTOp_(op, [x], [np.ones(3), np.ones(3)])
    assert g.tolist() == [5, 5, 5]
    op2 = O.par(O.op_scale(2.0), O.op_negate())
    assert [a.tolist() for a in O.gradTOp_(op2, [x, x], [x, x])] == [[2, 4, 6], [-1, -2, -3]]


def test_diag_getdiag_sumrows_maprows():
    v = np.array([1.0, 2.0, 3.0])
    d = O.diag(3, v)
    assert d.shape == (3, 3, 3) and d[1, 1, 1] == 2 and d.sum() == 6
    np.testing.assert_array_equal(O.getDiag(d), v)
    x = np.arange(6.0).reshape(2, 3)
    np.testing.assert_array_equal(O.sumRows(x), [3, 5, 7])
    (g,) = O.gradTOp_(O.op_sumRows(), [x], [v])
    np.testing.assert_array_equal(g, np.stack([v, v]))


def test_dots_two_circles_is_learned():
    """config 1 (app/Dots.hs): 2->16->1 logistic, squaredError, rate 1, per-sample SGD (reduced sample count)."""
    net = O.dots_train(n_samples=12000, hidden=(16,), rate=1.0, seed=0)
    assert O.dots_accuracy(net) > 0.85
