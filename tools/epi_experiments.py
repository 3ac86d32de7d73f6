"""GPU experiment: cost of each fused epilogue of the forward GEMM at config-2 size (per-kernel CUDA-event times)."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_ops_b200 as tb
from tensor_ops_b200 import nn, _lib as L

ctx = tb.Context(0)
B, i, o = 65536, 1024, 1024
X = ctx.rand_uniform((B, i), -1, 1, seed=1); dA = ctx.rand_normal((B, o), 0, 1, seed=2)
W = ctx.rand_normal((o, i), 0, 0.5, seed=3); b = ctx.rand_normal((o,), 0, 0.5, seed=4)
A = ctx.empty((B, o)); dX = ctx.empty((B, i)); dW = ctx.empty((o, i)); db = ctx.empty((o,))

def run(name, fn, n=10):
    for _ in range(3): fn()
    ctx.sync(); ctx.profile(True)
    for _ in range(n): fn()
    p = ctx.profile_summary(); ctx.profile(False)
    print(name, {k: round(v["ms"] / v["launches"], 4) for k, v in p.items()}, flush=True)

def fwd_only():
    s = L.c_buf(A.b.value); ctx.check(L.lib.tops_fflayer_fwd(ctx.h, X.b, W.b, b.b, 1, C.byref(s)))
def fg(with_db, with_dx=True):
    s = [L.c_buf(A.b.value), L.c_buf(dX.b.value), L.c_buf(dW.b.value), L.c_buf(db.b.value)]
    ctx.check(L.lib.tops_fflayer_fwd_grad(ctx.h, X.b, W.b, b.b, 1, dA.b, C.byref(s[0]), C.byref(s[1]) if with_dx else None, C.byref(s[2]),
                                          C.byref(s[3]) if with_db else None))
for prec, pn in ((tb.PREC_TF32, "tf32"), (tb.PREC_TF32X3, "tf32x3"), (tb.PREC_TF32_BF16X2, "tf32bf16")):
    ctx.set_precision(prec)
    run(pn + " fwd only (BIAS_ACT)", fwd_only)
    run(pn + " fwd_grad no db (DZ)", lambda: fg(False))
    run(pn + " fwd_grad + db (DZ+colsum)", lambda: fg(True))
