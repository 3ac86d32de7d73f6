"""GPU parity tests: the CUDA path, called through the C ABI (ctypes -> libtops_b200.so), against the fp64 CPU oracle
and the committed golden fixtures.

Tolerances (fp32 device vs fp64 oracle), norm-wise  ||dev - ref||_F / ||ref||_F  per output tensor:
    TOPS_PREC_F16X3 (default), TOPS_PREC_TF32_BF16X2, TOPS_PREC_TF32X3 and TOPS_PREC_FP32_SIMT : 1e-5   (BASELINE.json north_star: "within 1e-5 relative fp32")
    TOPS_PREC_TF32 (throughput mode, 10-bit mantissa)  : 3e-3   (NOT a parity mode; checked so it cannot silently rot)
    bf16 storage + fp32 accumulate                     : 2e-2 against the oracle run on bf16-rounded inputs
"""
import os

import numpy as np
import pytest

import tensor_ops_b200 as tb
from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import expr as E
from tensor_ops_b200 import nn, top as TO
from tensor_ops_b200.batched import BatchT

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = {tb.PREC_F16X3: 1e-5, tb.PREC_TF32_BF16X2: 1e-5, tb.PREC_TF32X3: 1e-5, tb.PREC_FP32_SIMT: 1e-5, tb.PREC_TF32: 3e-3}
PRECS = [tb.PREC_F16X3, tb.PREC_TF32_BF16X2, tb.PREC_TF32X3, tb.PREC_TF32, tb.PREC_FP32_SIMT]
DEFAULT_PREC = tb.PREC_F16X3


@pytest.fixture(scope="module")
def ctx():
    c = tb.Context(0)
    yield c
    c.set_precision(DEFAULT_PREC)


def rel(got, ref):
    got = np.asarray(got, dtype=np.float64); ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.isfinite(got).all()
    return float(np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-300))


def close(got, ref, tol, what=""):
    e = rel(got.numpy() if hasattr(got, "numpy") else got, ref)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"


# ------------------------------------------------------------------------------------------ fused ffLayer path
@pytest.mark.parametrize("prec", PRECS)
def test_fflayer_golden(ctx, prec):
    g = np.load(os.path.join(G, "fflayer_logistic.npz"))
    ctx.set_precision(prec)
    got = nn.fflayer_fwd_grad(*(ctx.from_numpy(g[k]) for k in ("X", "W", "b", "dA")))
    for name, t in zip(("A", "dX", "dW", "db"), got):
        close(t, g[name], TOL[prec], name)
    close(nn.fflayer_fwd(ctx.from_numpy(g["X"]), ctx.from_numpy(g["W"]), ctx.from_numpy(g["b"])), g["A"], TOL[prec], "fwd")
    ctx.set_precision(DEFAULT_PREC)


# empty, single-sample, ragged (not multiples of any tile), one-wide, and multi-tile shapes
SHAPES = [(1, 1, 1), (1, 7, 3), (5, 3, 1), (33, 129, 65), (256, 96, 72), (300, 260, 131), (1000, 520, 200), (2048, 1024, 1024),
          (4097, 257, 384)]


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,i,o", SHAPES)
def test_fflayer_fwd_grad_shapes(ctx, prec, B, i, o):
    ctx.set_precision(prec)
    rng = np.random.default_rng(B * 7 + i * 3 + o)
    X = rng.uniform(-1, 1, (B, i)); W = rng.normal(0, 0.5, (o, i)); b = rng.normal(0, 0.5, o); dA = rng.normal(size=(B, o))
    # the oracle sees exactly the fp32 values the device gets
    X, W, b, dA = (a.astype(np.float32).astype(np.float64) for a in (X, W, b, dA))
    ref = O.fflayer_logistic_dense(X, W, b, dA)
    dX, dW, db = ctx.from_numpy(X), ctx.from_numpy(W), ctx.from_numpy(b)
    got = nn.fflayer_fwd_grad(dX, dW, db, ctx.from_numpy(dA))
    for name, t, r in zip(("A", "dX", "dW", "db"), got, ref):
        close(t, r, TOL[prec], f"{name} B={B} i={i} o={o}")
    # gradTOp' alone: with recomputation (Types.hs:155) and from a saved activation
    g1 = nn.fflayer_grad(dX, dW, db, ctx.from_numpy(dA))
    g2 = nn.fflayer_grad(dX, dW, db, ctx.from_numpy(dA), A_saved=got[0])
    for name, a, c, r in zip(("dX", "dW", "db"), g1, g2, ref[1:]):
        close(a, r, TOL[prec], name + " (recompute)")
        close(c, r, TOL[prec], name + " (saved A)")
    ctx.set_precision(DEFAULT_PREC)


def test_fflayer_empty_batch(ctx):
    W, b = ctx.from_numpy(np.ones((4, 3))), ctx.from_numpy(np.zeros(4))
    X = ctx.empty((0, 3)); dA = ctx.empty((0, 4))
    A, dX, dW, db = nn.fflayer_fwd_grad(X, W, b, dA)
    assert A.shape == (0, 4) and dX.shape == (0, 3)
    assert np.all(dW.numpy() == 0) and np.all(db.numpy() == 0)


def test_fflayer_identity_activation_and_non_saturating_init(ctx):
    rng = np.random.default_rng(3)
    B, i, o = 512, 320, 192
    X = rng.uniform(-1, 1, (B, i)).astype(np.float32).astype(np.float64)
    W = rng.normal(0, 1 / np.sqrt(i), (o, i)).astype(np.float32).astype(np.float64)
    b = rng.normal(0, 0.5, o).astype(np.float32).astype(np.float64)
    dA = rng.normal(size=(B, o)).astype(np.float32).astype(np.float64)
    A, dX, dW, db = nn.fflayer_fwd_grad(ctx.from_numpy(X), ctx.from_numpy(W), ctx.from_numpy(b), ctx.from_numpy(dA), act=tb.ACT_ID)
    close(A, X @ W.T + b, 1e-5, "A"); close(dX, dA @ W, 1e-5, "dX"); close(dW, dA.T @ X, 1e-5, "dW"); close(db, dA.sum(0), 1e-5, "db")
    ref = O.fflayer_logistic_dense(X, W, b, dA)
    got = nn.fflayer_fwd_grad(ctx.from_numpy(X), ctx.from_numpy(W), ctx.from_numpy(b), ctx.from_numpy(dA))
    for name, t, r in zip(("A", "dX", "dW", "db"), got, ref):
        close(t, r, 1e-5, name)


def test_fflayer_degenerate_magnitudes(ctx):
    """The fp16-pair scaling at its corners: an all-zero cotangent (max|dA| = 0), all-zero rows of X, an all-zero W or X, and rows of X
    whose magnitudes differ by 2^20 (the per-row scales of X and the row factors of the cotangent pair share one budget: DESIGN §2).
    Parity must hold in all of them; with an absurd spread (rows 1e25 and 1e-25 times the others: beyond any fp16-based scheme) the
    outputs must at least stay finite."""
    rng = np.random.default_rng(123)
    B, i, o = 512, 256, 256
    f = lambda a: a.astype(np.float32).astype(np.float64)
    X = f(rng.uniform(-1, 1, (B, i))); X[3] = 0.0; X[200:210] = 0.0
    W = f(rng.normal(0, 0.05, (o, i))); b = f(rng.normal(0, 0.5, o)); dA = f(rng.normal(size=(B, o)))
    Xo = X.copy(); Xo[5] *= 1e6; Xo[100] *= 1e-6                     # one outlier row 2^20 above the rest, one 2^20 below
    Xa = X.copy(); Xa[7] *= 1e25; Xa[11] *= 1e-25
    cases = {"zero rows": (X, W, dA, True), "dA = 0": (X, W, np.zeros_like(dA), True), "W = 0": (X, np.zeros_like(W), dA, True),
             "X = 0": (np.zeros_like(X), W, dA, True), "outlier rows": (f(Xo), W, dA, True), "absurd spread": (f(Xa), W, dA, False)}
    for name, (x, w, da, parity) in cases.items():
        with np.errstate(over="ignore"):
            ref = O.fflayer_logistic_dense(x, w, b, da)
        got = nn.fflayer_fwd_grad(ctx.from_numpy(x), ctx.from_numpy(w), ctx.from_numpy(b), ctx.from_numpy(da))
        for nm, t, r in zip(("A", "dX", "dW", "db"), got, ref):
            g = t.numpy()
            assert np.isfinite(g).all(), f"{name}: {nm} has non-finite values"
            if not parity:
                continue
            if np.linalg.norm(r) == 0.0:
                assert np.abs(g).max() == 0.0, f"{name}: {nm} should be exactly zero"
            else:
                close(t, r, 1e-5, f"{name} {nm}")


def test_fflayer_error_behaviour(ctx):
    X = ctx.from_numpy(np.zeros((4, 3))); W = ctx.from_numpy(np.zeros((5, 2))); b = ctx.from_numpy(np.zeros(5))
    with pytest.raises(tb.TopsError) as ei:
        nn.fflayer_fwd(X, W, b)
    assert ei.value.code == 2 and "FeedForward.hs" in str(ei.value)
    # the context stays usable after an error
    assert nn.fflayer_fwd(X, ctx.from_numpy(np.zeros((5, 3))), b).shape == (4, 5)


def test_fflayer_preallocated_packed_outputs(ctx):
    """dW and db written into views of ONE packed buffer [dW‖db] (what the data-parallel all-reduce sends)."""
    rng = np.random.default_rng(9)
    B, i, o = 200, 64, 48
    X, W, b, dA = rng.uniform(-1, 1, (B, i)), rng.normal(0, 0.5, (o, i)), rng.normal(0, 0.5, o), rng.normal(size=(B, o))
    packed = ctx.full((o * i + o,), 7.0)
    dWv, dbv = packed.view(0, (o, i)), packed.view(o * i, (o,))
    nn.fflayer_fwd_grad(ctx.from_numpy(X), ctx.from_numpy(W), ctx.from_numpy(b), ctx.from_numpy(dA), out=(None, None, dWv, dbv))
    ref = O.fflayer_logistic_dense(*(a.astype(np.float32).astype(np.float64) for a in (X, W, b, dA)))
    p = packed.numpy()
    close(p[:o * i].reshape(o, i), ref[2], 1e-5, "packed dW"); close(p[o * i:], ref[3], 1e-5, "packed db")


def test_fflayer_bf16(ctx):
    import torch
    rng = np.random.default_rng(5)
    B, i, o = 1024, 512, 384
    tX = torch.tensor(rng.uniform(-1, 1, (B, i)), dtype=torch.bfloat16, device="cuda")
    tW = torch.tensor(rng.normal(0, 1 / np.sqrt(i), (o, i)), dtype=torch.bfloat16, device="cuda")
    tdA = torch.tensor(rng.normal(size=(B, o)), dtype=torch.bfloat16, device="cuda")
    b = rng.normal(0, 0.5, o).astype(np.float32)
    torch.cuda.synchronize()
    got = nn.fflayer_fwd_grad(ctx.wrap_torch(tX), ctx.wrap_torch(tW), ctx.from_numpy(b), ctx.wrap_torch(tdA))
    ref = O.fflayer_logistic_dense(tX.double().cpu().numpy(), tW.double().cpu().numpy(), b.astype(np.float64), tdA.double().cpu().numpy())
    for name, t, r in zip(("A", "dX", "dW", "db"), got, ref):
        close(t, r, 2e-2, "bf16 " + name)
    assert got[0].dtype == tb.BF16 and got[2].dtype == tb.F32


# ------------------------------------------------------------------------------------------ MLP / netGrad
@pytest.mark.parametrize("prec", PRECS)
def test_mlp_softmax_ce_golden(ctx, prec):
    g = np.load(os.path.join(G, "mlp_softmax_ce.npz"))
    ctx.set_precision(prec)
    Ws = [ctx.from_numpy(g[f"W{l}"]) for l in range(3)]; bs = [ctx.from_numpy(g[f"b{l}"]) for l in range(3)]
    A, L, dX, dWs, dbs = nn.mlp_fwd_grad(Ws, bs, [tb.ACT_LOGISTIC, tb.ACT_LOGISTIC, tb.ACT_SOFTMAX], tb.LOSS_CROSS_ENTROPY,
                                         ctx.from_numpy(g["X"]), ctx.from_numpy(g["Y"]))
    t = TOL[prec]
    close(A, g["A"], t, "A"); close(dX, g["dX"], t, "dX")
    assert abs(L.unScalar() - float(g["L"])) <= max(t, 2e-6) * abs(float(g["L"]))
    for l in range(3):
        close(dWs[l], g[f"dW{l}"], t, f"dW{l}"); close(dbs[l], g[f"db{l}"], t, f"db{l}")
    close(nn.mlp_fwd(Ws, bs, [tb.ACT_LOGISTIC, tb.ACT_LOGISTIC, tb.ACT_SOFTMAX], ctx.from_numpy(g["X"])), g["A"], t, "mlp_fwd")
    ctx.set_precision(DEFAULT_PREC)


@pytest.mark.parametrize("acts,loss", [(["logistic", "logistic", "softmax"], "crossEntropy"), (["logistic", "logistic"], "squaredError"),
                                       (["logistic", "softmax", "logistic"], "squaredError"), (["id", "logistic"], "crossEntropy")])
def test_mlp_heads_vs_oracle(ctx, acts, loss):
    rng = np.random.default_rng(21)
    dims = [100, 72, 40, 10][:len(acts) + 1]
    B = 300
    Ws = [rng.normal(0, 1 / np.sqrt(dims[l]), (dims[l + 1], dims[l])) for l in range(len(acts))]
    bs = [rng.normal(0, 0.5, dims[l + 1]) for l in range(len(acts))]
    X = rng.uniform(0, 1, (B, dims[0]))
    Y = np.eye(dims[-1])[rng.integers(0, dims[-1], B)] if loss == "crossEntropy" else rng.uniform(0, 1, (B, dims[-1]))
    f = lambda a: a.astype(np.float32).astype(np.float64)
    Ws, bs, X, Y = [f(w) for w in Ws], [f(b) for b in bs], f(X), f(Y)
    ref = O.mlp_dense_fwd_grad(X, Ws, bs, acts, loss, Y)
    amap = {"logistic": tb.ACT_LOGISTIC, "softmax": tb.ACT_SOFTMAX, "id": tb.ACT_ID}
    A, L, dX, dWs, dbs = nn.mlp_fwd_grad([ctx.from_numpy(w) for w in Ws], [ctx.from_numpy(b) for b in bs], [amap[a] for a in acts],
                                         tb.LOSS_CROSS_ENTROPY if loss == "crossEntropy" else tb.LOSS_SQUARED_ERROR,
                                         ctx.from_numpy(X), ctx.from_numpy(Y))
    close(A, ref[0], 1e-5, "A"); close(dX, ref[2], 1e-5, "dX")
    assert abs(L.unScalar() - ref[1]) <= 1e-5 * abs(ref[1])
    for l in range(len(acts)):
        close(dWs[l], ref[3][l], 1e-5, f"dW{l}"); close(dbs[l], ref[4][l], 1e-5, f"db{l}")


@pytest.mark.parametrize("B,hidden,classes", [(1000, 64, 10), (777, 37, 5), (4100, 256, 16), (130, 33, 1), (3, 8, 2), (2500, 300, 13)])
def test_mlp_output_layer_skinny_products(ctx, B, hidden, classes):
    """An output layer with <= 16 classes: its three products (forward N <= 16, dW M <= 16, dA K <= 16) take the streaming fp32
    kernels (aligned float4 and unaligned scalar variants, every count of thin columns); checked against the oracle in every
    parity mode's tolerance through both heads, and through the bare ffLayer gradient with the same shapes."""
    rng = np.random.default_rng(B + hidden + classes)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    dims = [48, hidden, classes]
    Ws = [f(rng.normal(0, 1 / np.sqrt(dims[l]), (dims[l + 1], dims[l]))) for l in range(2)]
    bs = [f(rng.normal(0, 0.5, dims[l + 1])) for l in range(2)]
    X = f(rng.uniform(0, 1, (B, dims[0])))
    for acts, loss in ((["logistic", "softmax"], "crossEntropy"), (["logistic", "logistic"], "squaredError"), (["logistic", "softmax"], "squaredError")):
        if classes == 1 and acts[-1] == "softmax":
            continue                          # softmax over one class is the constant 1: every gradient is exactly zero
        Y = f(np.eye(classes)[rng.integers(0, classes, B)]) if loss == "crossEntropy" else f(rng.uniform(0, 1, (B, classes)))
        ref = O.mlp_dense_fwd_grad(X, Ws, bs, acts, loss, Y)
        amap = {"logistic": tb.ACT_LOGISTIC, "softmax": tb.ACT_SOFTMAX}
        n0 = ctx.launch_count()
        A, L, dX, dWs, dbs = nn.mlp_fwd_grad([ctx.from_numpy(w) for w in Ws], [ctx.from_numpy(b) for b in bs], [amap[a] for a in acts],
                                             tb.LOSS_CROSS_ENTROPY if loss == "crossEntropy" else tb.LOSS_SQUARED_ERROR, ctx.from_numpy(X), ctx.from_numpy(Y))
        assert ctx.launch_count() > n0
        close(A, ref[0], 1e-5, "A"); close(dX, ref[2], 1e-5, "dX")
        assert abs(L.unScalar() - ref[1]) <= 1e-5 * abs(ref[1])
        for l in range(2):
            close(dWs[l], ref[3][l], 1e-5, f"dW{l}"); close(dbs[l], ref[4][l], 1e-5, f"db{l}")
    # the same thin shapes through ffLayer' >>> logistic with a supplied cotangent, and a plain product with a thin inner dimension
    H = f(rng.uniform(0, 1, (B, hidden))); dA = f(rng.normal(size=(B, classes)))
    ref = O.fflayer_logistic_dense(H, Ws[1], bs[1], dA)
    got = nn.fflayer_fwd_grad(ctx.from_numpy(H), ctx.from_numpy(Ws[1]), ctx.from_numpy(bs[1]), ctx.from_numpy(dA))
    for name, t, r in zip(("A", "dX", "dW", "db"), got, ref):
        close(t, r, 1e-5, f"fflayer {name}")
    from tensor_ops_b200.tensor import CuTensor
    close(ctx.from_numpy(dA).gemm(ctx.from_numpy(Ws[1])), dA @ Ws[1], 1e-5, "thin-K product")
    close(ctx.from_numpy(H).gemm(CuTensor.transp(ctx.from_numpy(Ws[1]))), H @ Ws[1].T, 1e-5, "thin-N product")


def test_narrow_output_layer_on_the_tcgen05_route():
    """TOPS_SKINNY=0 keeps products with a dimension <= 16 on the tensor-core path (fp16 planes with padded rows), the route they took
    before the streaming kernels existed: same netGrad, same tolerance."""
    os.environ["TOPS_SKINNY"] = "0"
    try:
        c2 = tb.Context(0)
    finally:
        del os.environ["TOPS_SKINNY"]
    rng = np.random.default_rng(77)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    dims, B = [48, 64, 10], 1000
    Ws = [f(rng.normal(0, 1 / np.sqrt(dims[l]), (dims[l + 1], dims[l]))) for l in range(2)]
    bs = [f(rng.normal(0, 0.5, dims[l + 1])) for l in range(2)]
    X = f(rng.uniform(0, 1, (B, dims[0]))); Y = f(np.eye(10)[rng.integers(0, 10, B)])
    ref = O.mlp_dense_fwd_grad(X, Ws, bs, ["logistic", "softmax"], "crossEntropy", Y)
    A, L, dX, dWs, dbs = nn.mlp_fwd_grad([c2.from_numpy(w) for w in Ws], [c2.from_numpy(b) for b in bs], [tb.ACT_LOGISTIC, tb.ACT_SOFTMAX],
                                         tb.LOSS_CROSS_ENTROPY, c2.from_numpy(X), c2.from_numpy(Y))
    close(A, ref[0], 1e-5, "A"); close(dX, ref[2], 1e-5, "dX")
    for l in range(2):
        close(dWs[l], ref[3][l], 1e-5, f"dW{l}"); close(dbs[l], ref[4][l], 1e-5, f"db{l}")


@pytest.mark.parametrize("classes", [20, 32, 40])
def test_softmax_ce_head_wider_than_the_thread_per_row_kernel(ctx, classes):
    """17..32 classes: warp-per-row head with db fused by lane; more than 32: the head leaves db to the column-sum pass."""
    rng = np.random.default_rng(classes)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    dims, B = [30, 24, classes], 515
    Ws = [f(rng.normal(0, 1 / np.sqrt(dims[l]), (dims[l + 1], dims[l]))) for l in range(2)]
    bs = [f(rng.normal(0, 0.5, dims[l + 1])) for l in range(2)]
    X = f(rng.uniform(0, 1, (B, dims[0]))); Y = f(np.eye(classes)[rng.integers(0, classes, B)])
    ref = O.mlp_dense_fwd_grad(X, Ws, bs, ["logistic", "softmax"], "crossEntropy", Y)
    A, L, dX, dWs, dbs = nn.mlp_fwd_grad([ctx.from_numpy(w) for w in Ws], [ctx.from_numpy(b) for b in bs], [tb.ACT_LOGISTIC, tb.ACT_SOFTMAX],
                                         tb.LOSS_CROSS_ENTROPY, ctx.from_numpy(X), ctx.from_numpy(Y))
    close(A, ref[0], 1e-5, "A"); close(dX, ref[2], 1e-5, "dX")
    assert abs(L.unScalar() - ref[1]) <= 1e-5 * abs(ref[1])
    for l in range(2):
        close(dWs[l], ref[3][l], 1e-5, f"dW{l}"); close(dbs[l], ref[4][l], 1e-5, f"db{l}")


# ------------------------------------------------------------------------------------------ per-sample TOp algebra on the device
def test_dots_golden_per_sample_netgrad_and_training(ctx):
    """config 1 plumbing (tensor-ops-dots 2->16->1): netGrad per sample and 200 per-sample SGD steps (Dots.hs:74-80)."""
    g = np.load(os.path.join(G, "dots_2_16_1.npz"))
    net = nn.networkFromParams([ctx.from_numpy(g[k]) for k in ("W0", "b0", "W1", "b1")], [nn.actLogistic, nn.actLogistic])
    for s in range(8):
        gr = nn.netGrad(nn.squaredError(), ctx.from_numpy(g["xs"][s]), ctx.from_numpy(g["ys"][s]), net)
        for name, t in zip(("dx", "dW0", "db0", "dW1", "db1"), gr):
            e = np.abs(t.numpy() - g[name][s]).max()
            assert e <= 1e-6 * max(1.0, np.abs(g[name][s]).max()), (name, s, e)
    n2 = net
    for s in range(200):
        n2 = nn.trainNetwork(nn.squaredError(), 1.0, ctx.from_numpy(g["xs"][s]), ctx.from_numpy(g["ys"][s]), n2)
    for name, p in zip(("W0_200", "b0_200", "W1_200", "b1_200"), n2.params):
        close(p, g[name], 2e-5, name)   # 200 sequential fp32 steps vs fp64


def test_generic_batched_instance_matches_fused_path(ctx):
    """The vmap-style BatchT `instance Tensor` (closures, no pattern matching) and the fused kernels agree."""
    rng = np.random.default_rng(17)
    B = 64
    Ws = [rng.normal(0, 0.3, (12, 20)), rng.normal(0, 0.3, (6, 12))]; bs = [rng.normal(0, 0.5, 12), rng.normal(0, 0.5, 6)]
    X = rng.uniform(-1, 1, (B, 20)); Y = rng.uniform(0, 1, (B, 6))
    params = [ctx.from_numpy(a) for a in (Ws[0], bs[0], Ws[1], bs[1])]
    net = nn.networkFromParams(params, [nn.actLogistic, nn.actLogistic])
    ls, dX, grads = nn.netGradBatched(nn.squaredError(), ctx.from_numpy(X), ctx.from_numpy(Y), net)
    opaque = nn.Network(net.op, net.params, None)      # forget the layer tags: forces the generic BatchT evaluation
    ls2, dX2, grads2 = nn.netGradBatched(nn.squaredError(), ctx.from_numpy(X), ctx.from_numpy(Y), opaque)
    ref = O.mlp_dense_fwd_grad(X, Ws, bs, ["logistic", "logistic"], "squaredError", Y)
    assert abs(ls.unScalar() - ref[1]) < 1e-4 * ref[1] and abs(ls2.unScalar() - ref[1]) < 1e-4 * ref[1]
    dX2 = dX2.t if isinstance(dX2, BatchT) else dX2
    close(dX, ref[2], 1e-5, "dX fused"); close(dX2, ref[2], 1e-5, "dX generic")
    want = [ref[3][0], ref[4][0], ref[3][1], ref[4][1]]
    for a, c, r in zip(grads, grads2, want):
        close(a, r, 1e-5, "fused"); close(c, r, 1e-5, "generic")
    close(nn.runNetworkBatched(opaque, ctx.from_numpy(X)), ref[0], 1e-5, "runNetworkBatched generic")


# ------------------------------------------------------------------------------------------ class BLAS / class Tensor methods
def test_blas_level1_and_2(ctx):
    rng = np.random.default_rng(2)
    x, y = rng.normal(size=1000), rng.normal(size=1000)
    dx, dy = ctx.from_numpy(x), ctx.from_numpy(y)
    close(dx.axpy(2.5, dy), 2.5 * x + y, 1e-6, "axpy"); close(dx.axpy(-3.0), -3.0 * x, 1e-6, "axpy Nothing")
    close(dx + dy, x + y, 1e-6, "add")
    assert abs(dx.dot(dy).unScalar() - x @ y) < 1e-4 * np.linalg.norm(x) * np.linalg.norm(y)
    assert abs(dx.sum().unScalar() - x.sum()) < 1e-4 * np.abs(x).sum()
    a = rng.normal(size=(37, 1000)); da = ctx.from_numpy(a)
    u = rng.normal(size=37); du = ctx.from_numpy(u)
    close(da.gemv(dx), a @ x, 1e-5, "gemv"); close(da.gemv(dx, 0.5, 2.0, du), 0.5 * a @ x + 2.0 * u, 1e-5, "gemv beta")
    close(tb.CuTensor.transp(da).gemv(du), a.T @ u, 1e-5, "gemv (tr A)")
    close(du.ger(dx), np.outer(u, x), 1e-6, "ger")
    assert abs(da.index((3, 5)) - np.float32(a[3, 5])) < 1e-7 and abs(tb.CuTensor.transp(da).index((5, 3)) - np.float32(a[3, 5])) < 1e-7
    close(da.row(4), a[4], 1e-7, "indexRowB")
    sq = rng.normal(size=(50, 50)); dsq = ctx.from_numpy(sq)
    assert abs(dsq.trace().unScalar() - np.trace(sq)) < 1e-4
    close(tb.tensor.eye(ctx, 9), np.eye(9), 0, "eye")
    close(tb.CuTensor.getDiag(dsq), np.diag(sq), 1e-7, "getDiag")
    close(tb.CuTensor.diag(2, du), np.diag(u), 1e-7, "diag")
    d3 = tb.CuTensor.diag(3, du).numpy()
    assert d3.shape == (37, 37, 37) and np.count_nonzero(d3) == 37 and abs(d3[5, 5, 5] - np.float32(u[5])) < 1e-7
    close(tb.CuTensor.sumRows(da), a.sum(0), 1e-5, "sumRows"); close(tb.CuTensor.broadcastRows(5, du), np.tile(u, (5, 1)), 1e-7, "broadcastRows")
    close(tb.CuTensor.sumT([da, da, da]), 3 * a, 1e-6, "sumT")


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("n,k,m", [(64, 64, 64), (130, 70, 33), (512, 300, 257), (1, 5, 9)])
def test_gemm_all_transposition_views(ctx, prec, n, k, m):
    """`gemm alpha A B (Just (beta, C))` (BLAS.hs:118-123) with O(1) `transpB` views as either operand."""
    ctx.set_precision(prec)
    rng = np.random.default_rng(n + k + m)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    a, b, c = f(rng.normal(size=(n, k))), f(rng.normal(size=(k, m))), f(rng.normal(size=(n, m)))
    da, db_, dc = ctx.from_numpy(a), ctx.from_numpy(b), ctx.from_numpy(c)
    dat, dbt = tb.CuTensor.transp(ctx.from_numpy(a.T.copy())), tb.CuTensor.transp(ctx.from_numpy(b.T.copy()))
    t = TOL[prec]
    close(da.gemm(db_), a @ b, t, "NN"); close(dat.gemm(db_), a @ b, t, "TN"); close(da.gemm(dbt), a @ b, t, "NT"); close(dat.gemm(dbt), a @ b, t, "TT")
    close(da.gemm(db_, 0.5, -2.0, dc), 0.5 * a @ b - 2.0 * c, t, "alpha/beta")
    ctx.set_precision(DEFAULT_PREC)


GMUL_CASES = [((), (5,), ()), ((4,), (), (3,)), ((4,), (6,), ()), ((), (6,), (5,)), ((4,), (6,), (5,)), ((3, 4), (5,), (6,)),
              ((3,), (4, 5), (6,)), ((3, 2), (4, 5), ()), ((2,), (3, 4, 2), (3,)), ((6, 5), (), (4,)), ((3,), (4,), (5, 2)),
              ((64, 64), (64,), (64,)), ((), (7, 3), ())]


@pytest.mark.parametrize("ms,os_,ns", GMUL_CASES)
def test_gmul_rank_dispatch_vs_two_oracle_witnesses(ctx, ms, os_, ns):
    rng = np.random.default_rng(len(ms) * 100 + len(os_) * 10 + len(ns))
    x = rng.normal(size=ms + os_).astype(np.float32).astype(np.float64)
    y = rng.normal(size=tuple(reversed(os_)) + ns).astype(np.float32).astype(np.float64)
    ref = O.gmul(len(ms), len(os_), len(ns), x, y)
    if x.size * y.size < 20000:
        np.testing.assert_allclose(O.gmul_naive(len(ms), len(os_), len(ns), x, y), ref, rtol=1e-10, atol=1e-12)
    got = tb.CuTensor.gmul(len(ms), len(os_), len(ns), ctx.from_numpy(x), ctx.from_numpy(y))
    assert got.shape == tuple(ms + ns)
    close(got, ref, 1e-5, f"gmul {ms} {os_} {ns}")


def test_gmul_rank3_golden_with_vjp(ctx):
    """config 5 interpretation at reduced size: inner (LS (LS LZ)) (LS LZ) >>> sumRows, forward and gradTOp'."""
    h = np.load(os.path.join(G, "gmul_rank3.npz"))
    op = TO.gmul(2, 1, 1) >> TO.sumRows()
    x, y, ct = (ctx.from_numpy(h[k]) for k in ("x", "y", "ct"))
    close(TO.runTOp(op, [x, y])[0], h["z"], 1e-5, "z")
    dx, dy = TO.gradTOp_(op, [x, y], [ct])
    close(dx, h["dx"], 1e-5, "dx"); close(dy, h["dy"], 1e-5, "dy")
    op2 = TO.gmul(1, 2, 1)
    x2, y2, ct2 = (ctx.from_numpy(h[k]) for k in ("x2", "y2", "ct2"))
    close(TO.runTOp(op2, [x2, y2])[0], h["z2"], 1e-5, "z2")
    dx2, dy2 = TO.gradTOp_(op2, [x2, y2], [ct2])
    close(dx2, h["dx2"], 1e-5, "dx2"); close(dy2, h["dy2"], 1e-5, "dy2")


def test_config5_contraction_full_size(ctx):
    """'[ '[64,64,64], '[64,64] ] -> '[ '[64,64] ]: fwd + gradTOp' with an explicit cotangent."""
    rng = np.random.default_rng(64)
    x, y, ct = (rng.normal(size=s).astype(np.float32).astype(np.float64) for s in ((64, 64, 64), (64, 64), (64, 64)))
    oop = O.op_gmul(2, 1, 1) >> O.op_sumRows()
    op = TO.gmul(2, 1, 1) >> TO.sumRows()
    dx_, dy_, dct = ctx.from_numpy(x), ctx.from_numpy(y), ctx.from_numpy(ct)
    close(TO.runTOp(op, [dx_, dy_])[0], O.runTOp(oop, [x, y])[0], 1e-5, "z")
    want = O.gradTOp_(oop, [x, y], [ct])
    got = TO.gradTOp_(op, [dx_, dy_], [dct])
    close(got[0], want[0], 1e-5, "dx"); close(got[1], want[1], 1e-5, "dy")


@pytest.mark.parametrize("A,R,K,N", [(5, 3, 7, 2), (9, 70, 33, 12), (4, 20, 96, 40), (17, 8, 5, 600), (3, 130, 64, 64), (2, 600, 16, 8)])
def test_fused_contraction_row_sum_shapes(ctx, A, R, K, N):
    """gmul (2,1,1) >>> sumRows on ragged shapes: more rows than columns of y and the reverse, N above and below the block size, row
    counts above it — forward (staged y) and the deterministic VJP (dy written once per element, no zeroing, no reds); the VJP
    is also bit-reproducible from call to call."""
    rng = np.random.default_rng(A * 1000 + R * 7 + K + N)
    x, y, ct = (rng.normal(size=s).astype(np.float32).astype(np.float64) for s in ((A, R, K), (K, N), (R, N)))
    oop = O.op_gmul(2, 1, 1) >> O.op_sumRows()
    op = TO.gmul(2, 1, 1) >> TO.sumRows()
    dx_, dy_, dct = ctx.from_numpy(x), ctx.from_numpy(y), ctx.from_numpy(ct)
    close(TO.runTOp(op, [dx_, dy_])[0], O.runTOp(oop, [x, y])[0], 1e-5, "z")
    want = O.gradTOp_(oop, [x, y], [ct])
    got = TO.gradTOp_(op, [dx_, dy_], [dct])
    close(got[0], want[0], 1e-5, "dx"); close(got[1], want[1], 1e-5, "dy")
    again = TO.gradTOp_(op, [dx_, dy_], [dct])
    assert np.array_equal(got[1].numpy(), again[1].numpy())


def test_transp_is_full_axis_reversal(ctx):
    rng = np.random.default_rng(8)
    for shape in [(7,), (5, 9), (3, 4, 5), (2, 3, 4, 5)]:
        a = rng.normal(size=shape)
        t = tb.CuTensor.transp(ctx.from_numpy(a))
        assert t.shape == tuple(reversed(shape))
        close(t, np.transpose(a), 1e-7, f"transp {shape}")


def test_lift_catalogue_and_interpreter(ctx):
    rng = np.random.default_rng(12)
    x = rng.normal(size=(33, 17)); y = rng.uniform(0.5, 2.0, size=(33, 17)); dx, dy = ctx.from_numpy(x), ctx.from_numpy(y)
    T = tb.CuTensor
    hits = tb._lib.lib.tops_lift_catalogue_hits
    h0 = hits()
    T.liftT(E.logistic, [dx]); T.liftT(lambda d, v: d * E.logistic_(v), [dy, dx]); T.liftT(E.exp, [dx]); T.liftT(E.log, [dy]); T.liftT(E.recip, [dy])
    T.liftT(lambda p, g: p - 0.02 * g, [dx, dy]); T.liftT(lambda a, b: (a - b) * (a - b), [dx, dy]); T.liftT(lambda a, b: a * b, [dx, dy])
    assert hits() - h0 == 8, "the reference's own lifted closures are all served by specialised kernels"
    T.liftT(lambda a, b: E.tanh(a) ** 2 + E.sqrt(b * b + 1) / (1 + a * a), [dx, dy])
    assert hits() - h0 == 8, "a program outside the catalogue runs in the interpreter"
    close(T.liftT(lambda a, b: (a - b) * (a - b), [dx, dy]), (x - y) ** 2, 1e-6, "(a-b)^2")
    close(T.liftT(lambda a, b: a * b, [dx, dy]), x * y, 1e-6, "a*b")
    close(T.liftT(E.logistic, [dx]), 1 / (1 + np.exp(-x)), 1e-6, "logistic")
    close(T.liftT(lambda d, v: d * E.logistic_(v), [dy, dx]), y * (lambda s: s * (1 - s))(1 / (1 + np.exp(-x))), 1e-6, "d*logistic'")
    close(T.liftT(E.exp, [dx]), np.exp(x), 1e-6, "exp"); close(T.liftT(E.log, [dy]), np.log(y), 1e-5, "log")
    close(T.liftT(E.recip, [dy]), 1 / y, 1e-6, "recip")
    close(T.liftT(lambda p, g: p - 0.02 * g, [dx, dy]), x - 0.02 * y, 1e-6, "sgd rule")
    close(T.liftT(lambda a, b: E.tanh(a) ** 2 + E.sqrt(b * b + 1) / (1 + a * a), [dx, dy]), np.tanh(x) ** 2 + np.sqrt(y * y + 1) / (1 + x * x), 1e-6, "generic")
    # map (derivative by symbolic differentiation, TOp.hs:207-213) through gradTOp'
    op = TO.map(lambda v: E.log(v * v + 3))
    (g,) = TO.gradTOp_(op, [dx], [dy])
    close(g, y * 2 * x / (x * x + 3), 1e-5, "map VJP")


def test_softmax_reference_form(ctx):
    rng = np.random.default_rng(14)
    z = rng.normal(size=(40, 10)); e = np.exp(z)
    zt = ctx.from_numpy(z)   # keep the handle alive: `.b` of a temporary would be released before the call
    close(zt._new(tb._lib.lib.tops_map_rows_softmax, zt.b), e / e.sum(1, keepdims=True), 1e-6, "softmax rows")
    for rows, cols in ((1000, 3), (517, 12), (4096, 16), (300, 20), (64, 100)):   # thread-per-row (scalar / 16-byte) and warp-per-row kernels
        z = rng.normal(size=(rows, cols)); e = np.exp(z); zt = ctx.from_numpy(z)
        close(zt._new(tb._lib.lib.tops_map_rows_softmax, zt.b), e / e.sum(1, keepdims=True), 1e-6, f"softmax rows {rows}x{cols}")
    # the reference softmax TOp on one sample, through the generic algebra
    v = rng.normal(size=10)
    close(TO.runTOp(nn.softmax(), [ctx.from_numpy(v)])[0], np.exp(v) / np.exp(v).sum(), 1e-6, "softmax TOp")


def test_device_rng_moments_and_determinism(ctx):
    a = ctx.rand_normal((1 << 20,), 0.0, 0.5, seed=3).numpy(); b = ctx.rand_normal((1 << 20,), 0.0, 0.5, seed=3).numpy()
    assert np.array_equal(a, b) and abs(a.mean()) < 3e-3 and abs(a.std() - 0.5) < 3e-3
    u = ctx.rand_uniform((1 << 20,), -1.0, 1.0, seed=4).numpy()
    assert u.min() >= -1 and u.max() <= 1 and abs(u.mean()) < 3e-3 and abs(u.std() - 2 / np.sqrt(12)) < 3e-3


def test_kernels_actually_launched(ctx):
    n0 = ctx.launch_count()
    nn.fflayer_fwd(ctx.from_numpy(np.ones((8, 8))), ctx.from_numpy(np.ones((8, 8))), ctx.from_numpy(np.ones(8)))
    assert ctx.launch_count() > n0


# ------------------------------------------------------------------------------------------ full BASELINE sizes
def test_config2_full_size_vs_oracle_and_properties(ctx):
    """ffLayer 1024->1024, batch 65536, fp32 (BASELINE.json configs[1]): whole-tensor comparison with the fp64 oracle
    plus size-independent properties (additivity of the parameter gradients over batch shards; db == column sums)."""
    rng = np.random.default_rng(0)
    B, i, o = 65536, 1024, 1024
    X = rng.uniform(-1, 1, (B, i)).astype(np.float32); W = rng.normal(0, 0.5, (o, i)).astype(np.float32)
    b = rng.normal(0, 0.5, o).astype(np.float32); dA = rng.standard_normal((B, o), dtype=np.float32)
    dX_, dW_, db_, ddA = ctx.from_numpy(X), ctx.from_numpy(W), ctx.from_numpy(b), ctx.from_numpy(dA)
    A, gX, gW, gb = nn.fflayer_fwd_grad(dX_, dW_, db_, ddA)
    ref = O.fflayer_logistic_dense(X.astype(np.float64), W.astype(np.float64), b.astype(np.float64), dA.astype(np.float64))
    for name, t, r in zip(("A", "dX", "dW", "db"), (A, gX, gW, gb), ref):
        close(t, r, 1e-5, f"config2 {name}")
    # additivity: gradients of the two half-batches sum to the full-batch gradient (what data-parallel sharding relies on)
    h = B // 2
    halves = []
    for r in range(2):
        Xh, dAh = dX_.view(r * h * i, (h, i)), ddA.view(r * h * o, (h, o))
        halves.append(nn.fflayer_fwd_grad(Xh, dW_, db_, dAh))
    close(halves[0][2].numpy().astype(np.float64) + halves[1][2].numpy(), ref[2], 1e-5, "dW additivity")
    close(halves[0][3].numpy().astype(np.float64) + halves[1][3].numpy(), ref[3], 1e-5, "db additivity")
    close(np.concatenate([halves[0][0].numpy(), halves[1][0].numpy()]), ref[0], 1e-5, "A shards")


def test_config3_mnist_shaped_mlp_full_batch(ctx):
    """3-layer MLP 784->512->256->10, logistic/logistic/softmax + crossEntropy, batch 32768 (configs[2], synthetic)."""
    rng = np.random.default_rng(1)
    B, dims = 32768, [784, 512, 256, 10]
    f = lambda a: a.astype(np.float32).astype(np.float64)
    Ws = [f(rng.normal(0, 1 / np.sqrt(dims[l]), (dims[l + 1], dims[l]))) for l in range(3)]
    bs = [f(rng.normal(0, 0.5, dims[l + 1])) for l in range(3)]
    X = f(rng.uniform(0, 1, (B, 784))); Y = np.eye(10)[rng.integers(0, 10, B)]
    ref = O.mlp_dense_fwd_grad(X, Ws, bs, ["logistic", "logistic", "softmax"], "crossEntropy", Y)
    A, L, dX, dWs, dbs = nn.mlp_fwd_grad([ctx.from_numpy(w) for w in Ws], [ctx.from_numpy(b) for b in bs],
                                         [tb.ACT_LOGISTIC, tb.ACT_LOGISTIC, tb.ACT_SOFTMAX], tb.LOSS_CROSS_ENTROPY,
                                         ctx.from_numpy(X), ctx.from_numpy(Y))
    close(A, ref[0], 1e-5, "A"); close(dX, ref[2], 1e-5, "dX")
    assert abs(L.unScalar() - ref[1]) <= 1e-5 * abs(ref[1])
    for l in range(3):
        close(dWs[l], ref[3][l], 1e-5, f"dW{l}"); close(dbs[l], ref[4][l], 1e-5, f"db{l}")


@pytest.mark.parametrize("B,n_chunks", [(1000, 3), (64, 0), (5000, 8)])
def test_host_buffer_entry_point_pipelined(ctx, B, n_chunks):
    """tops_fflayer_fwd_grad_host: host X/dA in, chunked H2D overlapped with compute, dW/db accumulated across chunks."""
    ctx.set_precision(DEFAULT_PREC)
    rng = np.random.default_rng(77 + B)
    i, o = 96, 72
    X = rng.uniform(-1, 1, (B, i)).astype(np.float32); dA = rng.normal(size=(B, o)).astype(np.float32)
    W = rng.normal(0, 0.5, (o, i)).astype(np.float32); b = rng.normal(0, 0.5, o).astype(np.float32)
    ref = O.fflayer_logistic_dense(X.astype(np.float64), W.astype(np.float64), b.astype(np.float64), dA.astype(np.float64))
    A = ctx.empty((B, o)); dX = ctx.empty((B, i)); packed = ctx.empty((o * i + o,))
    g = nn.fflayer_fwd_grad_host(ctx, X, ctx.from_numpy(W), ctx.from_numpy(b), dA, workspace=(A, dX, packed), n_chunks=n_chunks)
    close(A, ref[0], 1e-5, "A host path"); close(dX, ref[1], 1e-5, "dX host path")
    close(g[:o * i].reshape(o, i), ref[2], 1e-5, "dW host path"); close(g[o * i:], ref[3], 1e-5, "db host path")
    # the device copy of the packed gradient is what a data-parallel run all-reduces
    assert np.array_equal(packed.numpy(), g)


@pytest.mark.parametrize("prec,i,o", [(tb.PREC_F16X3, 784, 10), (tb.PREC_F16X3, 20, 1), (tb.PREC_TF32_BF16X2, 784, 10), (tb.PREC_FP32_SIMT, 96, 72),
                                      (tb.PREC_FP32_SIMT, 33, 10), (tb.PREC_F16X3, 1024, 1024)])
def test_host_buffer_entry_point_unfused_db_accumulates(ctx, prec, i, o):
    """when the forward epilogue cannot fuse the db column sums (o % 4 != 0 such as the MNIST head o = 10 or the Dots head
    o = 1, FP32_SIMT, the SIMT fallback) the host path must ACCUMULATE db across its row chunks instead of refusing."""
    ctx.set_precision(prec)
    rng = np.random.default_rng(i * 131 + o)
    B = 1500
    X = rng.uniform(-1, 1, (B, i)).astype(np.float32); dA = rng.normal(size=(B, o)).astype(np.float32)
    W = rng.normal(0, 0.5, (o, i)).astype(np.float32); b = rng.normal(0, 0.5, o).astype(np.float32)
    ref = O.fflayer_logistic_dense(X.astype(np.float64), W.astype(np.float64), b.astype(np.float64), dA.astype(np.float64))
    A = ctx.empty((B, o)); dX = ctx.empty((B, i)); packed = ctx.empty((o * i + o,))
    g = nn.fflayer_fwd_grad_host(ctx, X, ctx.from_numpy(W), ctx.from_numpy(b), dA, workspace=(A, dX, packed), n_chunks=4)
    close(A, ref[0], 1e-5, "A"); close(dX, ref[1], 1e-5, "dX")
    close(g[:o * i].reshape(o, i), ref[2], 1e-5, "dW"); close(g[o * i:], ref[3], 1e-5, "db accumulated over chunks")
    ctx.set_precision(DEFAULT_PREC)


@pytest.mark.parametrize("prec", [tb.PREC_F16X3, tb.PREC_TF32_BF16X2, tb.PREC_TF32X3, tb.PREC_TF32])
@pytest.mark.parametrize("B", [129, 255, 257, 385, 512])
def test_cta_pair_tile_edges(ctx, prec, B):
    """Row counts around the 128/256-row boundaries of the CTA-pair (cta_group::2) tiles: the peer CTA of the last pair owns
    1, 127, 1 (+ a full pair), ... valid rows; o = 264 also makes dW's own M straddle a pair boundary."""
    ctx.set_precision(prec)
    rng = np.random.default_rng(B)
    i, o = 200, 264
    X = rng.uniform(-1, 1, (B, i)); W = rng.normal(0, 0.5, (o, i)); b = rng.normal(0, 0.5, o); dA = rng.normal(size=(B, o))
    X, W, b, dA = (a.astype(np.float32).astype(np.float64) for a in (X, W, b, dA))
    ref = O.fflayer_logistic_dense(X, W, b, dA)
    got = nn.fflayer_fwd_grad(ctx.from_numpy(X), ctx.from_numpy(W), ctx.from_numpy(b), ctx.from_numpy(dA))
    for name, t, r in zip(("A", "dX", "dW", "db"), got, ref):
        close(t, r, TOL[prec], f"{name} B={B}")


def test_pinned_host_staging_buffers(ctx):
    """tops_host_alloc: page-locked (optionally write-combined) staging memory for upload / the host-buffer entry point."""
    rng = np.random.default_rng(3)
    for wc in (False, True):
        h = ctx.host_empty((37, 24), write_combined=wc)
        src = rng.standard_normal((37, 24)).astype(np.float32)
        h[...] = src
        t = ctx.empty((37, 24)).upload(h, sync=True)
        assert np.array_equal(t.numpy(), src)
