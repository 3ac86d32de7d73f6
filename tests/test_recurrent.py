"""Recurrent networks (src/TensorOps/Learn/NeuralNet/Recurrent.hs): the oracle's unroll/rollup BPTT against finite differences and
against a hand-written time loop; the product's host algebra against the oracle on CPU; the device against the oracle on GPU."""
import numpy as np
import pytest

from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import nn, recurrent as R, top as TO
from nptensor import NpT


def _oracle_rnn(rng, i=3, h=4, o=2):
    """fullyConnected(i->h, logistic state) ~*~ (ffLayer(h->o) *~ logistic): one recurrent layer feeding a stateless one."""
    l1 = O.r_fullyConnected(i, h, O.actLogistic, rng)
    l2 = O.r_then_act(O.r_stateless(O.ffLayer(h, o, rng)), O.actLogistic())
    return O.r_compose(l1, l2)


def _manual_sequence_loss(n, xs, ys):
    """Σ_t ||y_t - net(x_t | state)||² with the state threaded by hand through r_runNetwork."""
    tot, cur = 0.0, n
    for x, y in zip(xs, ys):
        out, cur = O.r_runNetwork(cur, x)
        tot += float(((y - out) ** 2).sum())
    return tot


def test_oracle_bptt_matches_manual_loop_and_finite_differences():
    rng = np.random.default_rng(21)
    n = _oracle_rnn(rng)
    xs = [rng.uniform(-1, 1, 3) for _ in range(4)]; ys = [rng.uniform(0, 1, 2) for _ in range(4)]
    gI, gS, gP = O.r_netGrad(O.squaredError(), xs, ys, n)
    assert len(gI) == 4 and len(gS) == 1 and len(gP) == 5
    base = _manual_sequence_loss(n, xs, ys)
    # the unrolled TOp computes the same total loss as the hand-written time loop
    T, nS, nP = 4, 1, 5
    full = O.firstOp(O.r_unroll(nS, nP, n.op, T) >> O.op_drop(nS, nS + T), T) >> O.r_rollup(O.squaredError(), T)
    assert abs(float(O.runTOp(full, xs[::-1] + n.state + n.params + ys)[0]) - base) < 1e-12
    eps = 1e-6
    def fd(mutate):
        up = mutate(+eps); dn = mutate(-eps)
        return (up - dn) / (2 * eps)
    # parameter W' (index 0), initial state, and the input of time step 1 (gI is in reversed time order: gI[3] belongs to xs[0])
    def mut_param(k, idx):
        def m(d):
            ps = [p.copy() for p in n.params]; ps[k][idx] += d
            return _manual_sequence_loss(O.RNetwork(n.op, n.state, ps), xs, ys)
        return m
    def mut_state(idx):
        def m(d):
            st = [s.copy() for s in n.state]; st[0][idx] += d
            return _manual_sequence_loss(O.RNetwork(n.op, st, n.params), xs, ys)
        return m
    def mut_input(t, idx):
        def m(d):
            x2 = [x.copy() for x in xs]; x2[t][idx] += d
            return _manual_sequence_loss(n, x2, ys)
        return m
    assert abs(fd(mut_param(0, (1, 2))) - gP[0][1, 2]) < 1e-7
    assert abs(fd(mut_param(1, (3, 0))) - gP[1][3, 0]) < 1e-7
    assert abs(fd(mut_param(4, (1,))) - gP[4][1]) < 1e-7
    assert abs(fd(mut_state(2)) - gS[0][2]) < 1e-7
    assert abs(fd(mut_input(0, 1)) - gI[3][1]) < 1e-7      # reversed order, as in the reference
    assert abs(fd(mut_input(3, 0)) - gI[0][0]) < 1e-7


def test_oracle_rnn_training_reduces_sequence_loss():
    rng = np.random.default_rng(2)
    n = _oracle_rnn(rng)
    xs = [rng.uniform(-1, 1, 3) for _ in range(5)]; ys = [np.array([0.2, 0.8]) for _ in range(5)]
    before = _manual_sequence_loss(n, xs, ys)
    for _ in range(200):
        n = O.r_trainNetwork(O.squaredError(), 0.1, 0.1, xs, ys, n)
    assert _manual_sequence_loss(n, xs, ys) < 0.3 * before


def _product_rnn(on, wrap):
    l1 = R.Network(R.fullyConnected_(nn.actLogistic), [wrap(on.state[0])], [wrap(p) for p in on.params[:3]])
    ff = nn.Network(nn.ffLayer_(), [wrap(p) for p in on.params[3:]], None)
    l2 = R.then_act(R.stateless(ff), nn.actLogistic.op())
    return R.compose(l1, l2)


def test_product_algebra_matches_oracle_on_numpy_tensor():
    rng = np.random.default_rng(8)
    on = _oracle_rnn(rng)
    xs = [rng.uniform(-1, 1, 3) for _ in range(3)]; ys = [rng.uniform(0, 1, 2) for _ in range(3)]
    pn = _product_rnn(on, NpT)
    gI, gS, gP = R.netGrad(nn.squaredError(), [NpT(x) for x in xs], [NpT(y) for y in ys], pn, NpT)
    wI, wS, wP = O.r_netGrad(O.squaredError(), xs, ys, on)
    for g, w in zip(gI + gS + gP, wI + wS + wP):
        np.testing.assert_allclose(g.a, w, rtol=1e-10, atol=1e-13)
    y0, pn2 = R.runNetwork(pn, NpT(xs[0]), NpT)
    w0, on2 = O.r_runNetwork(on, xs[0])
    np.testing.assert_allclose(y0.a, w0, rtol=1e-12)
    np.testing.assert_allclose(pn2.state[0].a, on2.state[0], rtol=1e-12)


@pytest.mark.gpu
def test_device_bptt_vs_oracle():
    import tensor_ops_b200 as tb
    ctx = tb.Context(0)
    rng = np.random.default_rng(13)
    on = _oracle_rnn(rng, i=24, h=40, o=8)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    on = O.RNetwork(on.op, [f32(s) for s in on.state], [f32(p) for p in on.params])
    xs = [f32(rng.uniform(-1, 1, 24)) for _ in range(6)]; ys = [f32(rng.uniform(0, 1, 8)) for _ in range(6)]
    pn = _product_rnn(on, ctx.from_numpy)
    gI, gS, gP = R.netGrad(nn.squaredError(), [ctx.from_numpy(x) for x in xs], [ctx.from_numpy(y) for y in ys], pn)
    wI, wS, wP = O.r_netGrad(O.squaredError(), xs, ys, on)
    for g, w in zip(gI + gS + gP, wI + wS + wP):
        assert np.linalg.norm(g.numpy().astype(np.float64) - w) <= 1e-5 * np.linalg.norm(w)
    n2 = R.trainNetwork(nn.squaredError(), 0.05, 0.1, [ctx.from_numpy(x) for x in xs], [ctx.from_numpy(y) for y in ys], pn)
    o2 = O.r_trainNetwork(O.squaredError(), 0.05, 0.1, xs, ys, on)
    for g, w in zip(n2.state + n2.params, o2.state + o2.params):
        assert np.linalg.norm(g.numpy().astype(np.float64) - w) <= 1e-5 * np.linalg.norm(w)
