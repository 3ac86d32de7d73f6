"""AutoEncoder (src/TensorOps/Learn/NeuralNet/AutoEncoder.hs): oracle self-checks, product host algebra vs oracle on CPU,
device vs oracle on the GPU (per-sample TOp algebra and the fused batched path)."""
import numpy as np
import pytest

from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import autoencoder as AE, nn, top as TO
from nptensor import NpT


def _oracle_encoder(rng, i=7, h=4):
    enc = O.genNet(i, [(5, O.actLogistic)], h, O.actLogistic, rng)
    dec = O.genNet(h, [], i, O.actLogistic, rng)
    return O.Encoder(enc, dec)


def test_oracle_encgrad_matches_finite_differences_and_netgrad():
    rng = np.random.default_rng(11)
    e = _oracle_encoder(rng)
    x = rng.uniform(0, 1, 7)
    gE, gD = O.encGrad(O.squaredError(), x, e)
    # the target is the input itself: same parameter gradients as netGrad loss x x (e >>> d)
    want = O.netGrad(O.squaredError(), x, x, O.encoderNet(e))[1:]
    for g, w in zip(gE + gD, want):
        np.testing.assert_allclose(g, w, rtol=1e-12, atol=1e-14)
    # central finite differences on one encoder weight and one decoder bias
    def loss_with(enc_params, dec_params):
        return O.testEncoder(O.squaredError(), O.Encoder(O.Network(e.enc.op, enc_params), O.Network(e.dec.op, dec_params)), x)
    eps = 1e-6
    for (which, k, idx) in (("enc", 0, (2, 3)), ("dec", 1, (4,))):
        ps = [p.copy() for p in (e.enc.params if which == "enc" else e.dec.params)]
        ps[k][idx] += eps
        up = loss_with(ps, e.dec.params) if which == "enc" else loss_with(e.enc.params, ps)
        ps[k][idx] -= 2 * eps
        dn = loss_with(ps, e.dec.params) if which == "enc" else loss_with(e.enc.params, ps)
        g = (gE if which == "enc" else gD)[k][idx]
        assert abs((up - dn) / (2 * eps) - g) < 1e-7 * max(1.0, abs(g))
    assert np.allclose(O.encodeDecode(e, x), O.decode(e, O.encode(e, x)))


def test_oracle_training_reduces_reconstruction_loss():
    rng = np.random.default_rng(3)
    e = _oracle_encoder(rng)
    xs = rng.uniform(0.2, 0.8, (300, 7))
    before = np.mean([O.testEncoder(O.squaredError(), e, x) for x in xs[:50]])
    for x in xs:
        e = O.trainEncoder(O.squaredError(), 0.5, x, e)
    after = np.mean([O.testEncoder(O.squaredError(), e, x) for x in xs[:50]])
    assert after < 0.7 * before


def _product_encoder(e, wrap):
    enc = nn.networkFromParams([wrap(p) for p in e.enc.params], [nn.actLogistic, nn.actLogistic])
    dec = nn.networkFromParams([wrap(p) for p in e.dec.params], [nn.actLogistic])
    return AE.Encoder(enc, dec)


def test_product_algebra_matches_oracle_on_numpy_tensor():
    rng = np.random.default_rng(5)
    e = _oracle_encoder(rng)
    x = rng.uniform(0, 1, 7)
    pe = _product_encoder(e, NpT)
    gE, gD = AE.encGrad(nn.squaredError(), NpT(x), pe, NpT)
    wE, wD = O.encGrad(O.squaredError(), x, e)
    for g, w in zip(gE + gD, wE + wD):
        np.testing.assert_allclose(g.a, w, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(AE.testEncoder(nn.squaredError(), pe, NpT(x), NpT).a, O.testEncoder(O.squaredError(), e, x), rtol=1e-12)


@pytest.mark.gpu
def test_device_encoder_per_sample_and_batched_vs_oracle():
    import tensor_ops_b200 as tb
    ctx = tb.Context(0)
    rng = np.random.default_rng(9)
    e = _oracle_encoder(rng, i=40, h=12)
    f32 = lambda a: np.asarray(a, np.float32).astype(np.float64)
    e = O.Encoder(O.Network(e.enc.op, [f32(p) for p in e.enc.params]), O.Network(e.dec.op, [f32(p) for p in e.dec.params]))
    X = f32(rng.uniform(0, 1, (300, 40)))
    pe = _product_encoder(e, ctx.from_numpy)
    rel = lambda g, w: np.linalg.norm(g.numpy().astype(np.float64) - w) / max(np.linalg.norm(w), 1e-30)
    # per sample, through the TOp algebra on the device
    gE, gD = AE.encGrad(nn.squaredError(), ctx.from_numpy(X[0]), pe)
    wE, wD = O.encGrad(O.squaredError(), X[0], e)
    for g, w in zip(gE + gD, wE + wD):
        assert rel(g, w) < 1e-5
    # batched: fused MLP path with the batch as its own target, gradients summed over samples
    ls, bE, bD = AE.encGradBatched(nn.squaredError(), ctx.from_numpy(X), pe)
    sums = None
    for x in X:
        a, b = O.encGrad(O.squaredError(), x, e)
        sums = [s + g for s, g in zip(sums, a + b)] if sums else list(a + b)
    for g, w in zip(bE + bD, sums):
        assert rel(g, w) < 1e-5
    want_loss = sum(O.testEncoder(O.squaredError(), e, x) for x in X)
    assert abs(float(ls.numpy()) - want_loss) < 1e-4 * want_loss
    # one batched training step moves the parameters exactly as p - r * summed gradient
    e2 = AE.trainEncoderBatched(nn.squaredError(), 0.01, ctx.from_numpy(X), pe)
    for p2, p, g in zip(e2.enc.params + e2.dec.params, e.enc.params + e.dec.params, sums):
        assert rel(p2, p - 0.01 * g) < 1e-5
