"""GPU parity at the sizes the small tests do not reach:
  * BASELINE configs[3] shapes: bf16 storage, 4096 -> 4096, 32768 rows (= one GPU's shard of the 262144-row batch at N = 8),
    against the fp64 oracle evaluated on the bf16-rounded inputs;
  * fp32 with K = 4096 (config-4 width) in both fp32-grade tensor-core modes — longer reductions than config 2's K = 1024;
  * two ranks on two GPUs: the fused NVLS all-reduce (tops_fflayer_fwd_grad_mc) and the overlapped NCCL schedule
    (tops_fflayer_step_dp) against plain NCCL and against the oracle's full-batch gradient (skipped with < 2 devices).
The oracle's big fp64 GEMMs run on the host cores: ~10-20 s per test."""
import os
import subprocess
import sys

import numpy as np
import pytest

import tensor_ops_b200 as tb
from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import nn

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(got, ref):
    got = np.asarray(got, dtype=np.float64)
    assert got.shape == ref.shape and np.isfinite(got).all()
    return float(np.linalg.norm(got - ref) / np.linalg.norm(ref))


@pytest.fixture(scope="module")
def ctx():
    c = tb.Context(0)
    yield c
    c.set_precision(tb.PREC_F16X3)


def test_config4_shapes_bf16_vs_oracle_on_rounded_inputs(ctx):
    import torch
    B, n = 32768, 4096
    g = torch.Generator(device="cuda").manual_seed(4)
    tX = (torch.rand((B, n), device="cuda", generator=g) * 2 - 1).to(torch.bfloat16)
    tdA = torch.randn((B, n), device="cuda", generator=g).to(torch.bfloat16)
    tW = (torch.randn((n, n), device="cuda", generator=g) * 0.5).to(torch.bfloat16)        # reference init N(0, 0.5^2), FeedForward.hs:206
    b = (np.random.default_rng(4).normal(0, 0.5, n)).astype(np.float32)
    torch.cuda.synchronize()
    A, dX, dW, db = nn.fflayer_fwd_grad(ctx.wrap_torch(tX), ctx.wrap_torch(tW), ctx.from_numpy(b), ctx.wrap_torch(tdA))
    assert A.dtype == tb.BF16 and dX.dtype == tb.BF16 and dW.dtype == tb.F32
    Xh, Wh, dAh = (t.double().cpu().numpy() for t in (tX, tW, tdA))
    rA, rdX, rdW, rdb = O.fflayer_logistic_dense(Xh, Wh, b.astype(np.float64), dAh)
    # bf16 outputs carry 2^-9 rounding (3.9e-3 max, ~1.1e-3 rms); dZ is stored in bf16 before the gradient GEMMs, which shows in dW/db/dX
    assert rel(A.numpy(), rA) < 4e-3
    assert rel(dX.numpy(), rdX) < 8e-3
    assert rel(dW.numpy(), rdW) < 8e-3
    assert rel(db.numpy(), rdb) < 8e-3


@pytest.mark.parametrize("prec", [tb.PREC_F16X3, tb.PREC_TF32_BF16X2])
@pytest.mark.parametrize("init", ["reference", "scaled"])
def test_fp32_k4096_vs_oracle(ctx, prec, init):
    """fp32 parity bar (1e-5) with 4096-long reductions; 'reference' init saturates the logistic (pre-activation sigma ~ 18), which
    amplifies the forward GEMM's error ~4x in the gradients.  The default mode F16X3 meets 1e-5 everywhere (measured 2-3e-6).
    TF32_BF16X2 (round 1's default) does NOT at this width with the saturating init — measured dX 1.16e-5 — which is one reason it
    was replaced; it is held to 3e-5 there so that the limitation stays visible instead of silently growing."""
    ctx.set_precision(prec)
    rng = np.random.default_rng(11)
    B, n = 4096, 4096
    X = rng.uniform(-1, 1, (B, n)).astype(np.float32)
    W = rng.normal(0, 0.5 if init == "reference" else 1 / np.sqrt(n), (n, n)).astype(np.float32)
    b = rng.normal(0, 0.5, n).astype(np.float32); dA = rng.normal(size=(B, n)).astype(np.float32)
    got = nn.fflayer_fwd_grad(*(ctx.from_numpy(a) for a in (X, W, b, dA)))
    ref = O.fflayer_logistic_dense(*(a.astype(np.float64) for a in (X, W, b, dA)))
    tol = 3e-5 if (prec == tb.PREC_TF32_BF16X2 and init == "reference") else 1e-5
    errs = {name: rel(t.numpy(), r) for name, t, r in zip(("A", "dX", "dW", "db"), got, ref)}
    print(f"K=4096 prec={prec} init={init}: {errs}")
    ctx.set_precision(tb.PREC_F16X3)
    for name, e in errs.items():
        assert e <= tol, f"{name} K=4096 init={init}: rel err {e:.3e} > {tol:.0e}"


def test_two_rank_data_parallel_paths_agree():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29611", os.path.join(ROOT, "tests", "dp_gpu_worker.py")], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, (out.stdout[-3000:], out.stderr[-3000:])
    assert "DP_GPU_WORKER_OK" in out.stdout
