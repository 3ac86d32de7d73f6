"""Deferred / recorded evaluation: a composed TOp's forward + reverse sweep recorded once (tops_graph_*) and replayed
as one CUDA graph launch; saved activations make the number of kernels linear in the depth of the composition; the per-sample
training fold of config 1 (app/Dots.hs:74-80) runs on the device backend as one graph replay per sample."""
import numpy as np
import pytest

import tensor_ops_b200 as tb
from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import nn, recurrent as R, top as TO

pytestmark = pytest.mark.gpu


def rel(got, ref):
    got = np.asarray(got, dtype=np.float64)
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))


@pytest.fixture(scope="module")
def ctx():
    return tb.Context(0)


def test_recorded_rank3_contraction_replays_on_new_inputs(ctx):
    """BASELINE configs[4] (inner (LS (LS LZ)) (LS LZ) >>> sumRows, forward + VJP) recorded once, replayed on fresh inputs."""
    rng = np.random.default_rng(5)
    op = TO.compose(TO.sumRows(), TO.inner(2, 1))
    oop = O.op_gmul(2, 1, 1) >> O.op_sumRows()
    x, y, d = ctx.empty((64, 64, 64)), ctx.empty((64, 64)), ctx.empty((64, 64))
    hx, hy, hd = (rng.normal(size=s).astype(np.float32) for s in ((64, 64, 64), (64, 64), (64, 64)))
    x.upload(hx); y.upload(hy); d.upload(hd, sync=True)
    with ctx.record() as g:
        (out,) = TO.runTOp(op, [x, y])
        dx, dy = TO.gradTOp_(op, [x, y], [d])
    assert 1 <= g.kernel_count() <= 3          # forward fused to one kernel, VJP to one (tops_gmul_sum_rows*)
    for trial in range(3):
        if trial:   # new values in the SAME input tensors, no re-recording
            hx, hy, hd = (rng.normal(size=s).astype(np.float32) for s in ((64, 64, 64), (64, 64), (64, 64)))
            x.upload(hx); y.upload(hy); d.upload(hd, sync=True)
        n0 = ctx.launch_count()
        g.launch()
        assert ctx.launch_count() - n0 == g.kernel_count()
        X, Y, D = (a.astype(np.float64) for a in (hx, hy, hd))
        assert rel(out.numpy(), O.runTOp(oop, [X, Y])[0]) <= 1e-5
        rdx, rdy = O.gradTOp_(oop, [X, Y], [D])
        assert rel(dx.numpy(), rdx) <= 1e-5 and rel(dy.numpy(), rdy) <= 1e-5
    g.close()


def test_recording_refuses_host_readbacks(ctx):
    x = ctx.from_numpy(np.ones(4, np.float32))
    with pytest.raises(tb.TopsError):
        with ctx.record():
            x.numpy()
    # the context is usable again afterwards
    assert np.array_equal(tb.CuTensor.scaleT(2.0, x).numpy(), 2 * np.ones(4, np.float32))


def _rnn(ctx, rng, i, h, o):
    l1 = O.r_fullyConnected(i, h, O.actLogistic, rng)
    l2 = O.r_then_act(O.r_stateless(O.ffLayer(h, o, rng)), O.actLogistic())
    on = O.r_compose(l1, l2)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    on = O.RNetwork(on.op, [f32(s) for s in on.state], [f32(p) for p in on.params])
    pl1 = R.Network(R.fullyConnected_(nn.actLogistic), [ctx.from_numpy(s) for s in on.state], [ctx.from_numpy(p) for p in on.params[:3]])
    pl2 = R.then_act(R.stateless(nn.Network(nn.ffLayer_(), [ctx.from_numpy(p) for p in on.params[3:]], None)), nn.actLogistic.op())
    return on, R.compose(pl1, pl2)


def test_bptt_kernel_count_is_linear_in_the_unrolled_length(ctx):
    """Recurrent.hs:392-431 `unroll` nests one `>>>` per time step: with the reference's chain rule (forward of the prefix re-run
    inside every gradient, Types.hs:155) a T-step BPTT launches O(T^2) kernels; with saved activations it is O(T)."""
    rng = np.random.default_rng(31)
    counts = {}
    for T in (8, 32):
        on, pn = _rnn(ctx, rng, 12, 16, 4)
        xs = [rng.uniform(-1, 1, 12).astype(np.float32) for _ in range(T)]; ys = [rng.uniform(0, 1, 4).astype(np.float32) for _ in range(T)]
        dxs, dys = [ctx.from_numpy(x) for x in xs], [ctx.from_numpy(y) for y in ys]
        n0 = ctx.launch_count()
        gI, gS, gP = R.netGrad(nn.squaredError(), dxs, dys, pn)
        counts[T] = ctx.launch_count() - n0
        wI, wS, wP = O.r_netGrad(O.squaredError(), [x.astype(np.float64) for x in xs], [y.astype(np.float64) for y in ys], on)
        for gg, w in zip(gI + gS + gP, wI + wS + wP):
            assert rel(gg.numpy(), w) <= 2e-5
    assert counts[32] <= 4.6 * counts[8], counts          # linear (4x) with slack, not quadratic (16x)


def test_dots_per_sample_fold_as_graph_replays(ctx):
    """config 1 (tensor-ops-dots 2->16->1, rate 1, squaredError, 50 000 samples) on the device backend: ONE recorded step
    (per-sample netGrad through the generic TOp evaluator + SGD update published in place), replayed once per sample."""
    rng = np.random.default_rng(0)
    n_samples = 50000
    inps = rng.uniform(-1, 1, (n_samples, 2)).astype(np.float32)
    outs = np.array([[O.dots_target(v.astype(np.float64))] for v in inps], dtype=np.float32)
    net = nn.genNet(2, [(16, nn.actLogistic)], 1, nn.actLogistic, seed=3, ctx=ctx)
    X_all, Y_all = ctx.from_numpy(inps), ctx.from_numpy(outs)
    x, y = ctx.empty((2,)), ctx.empty((1,))
    x.copy_from(X_all.row(0)); y.copy_from(Y_all.row(0))
    with ctx.record() as g:
        new = nn.trainNetwork(nn.squaredError(), 1.0, x, y, net)
        for p, q in zip(net.params, new.params):
            p.copy_from(q)
    per_step = g.kernel_count()
    for s in range(n_samples):
        x.copy_from(X_all.row(s)); y.copy_from(Y_all.row(s))
        g.launch()
    ctx.sync()
    pts = rng.uniform(-1, 1, (4000, 2)).astype(np.float32)
    pred = nn.runNetworkBatched(net, ctx.from_numpy(pts)).numpy()[:, 0] > 0.5
    want = np.array([O.dots_target(v.astype(np.float64)) > 0.5 for v in pts])
    acc = float((pred == want).mean())
    g.close()
    assert acc > 0.85, (acc, per_step)
