"""Launch a few config-2 fwd+grad steps (for ncu captures).  usage: one_fwd_grad.py [tf32|tf32x3] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tensor_ops_b200 as tb
from tensor_ops_b200 import nn
prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
ctx = tb.Context(0); ctx.set_precision({"tf32": tb.PREC_TF32, "tf32x3": tb.PREC_TF32X3, "tf32bf16": tb.PREC_TF32_BF16X2}[prec])
B, i, o = 65536, 1024, 1024
X = ctx.rand_uniform((B, i), -1, 1, seed=1); dA = ctx.rand_normal((B, o), 0, 1, seed=2)
W = ctx.rand_normal((o, i), 0, 0.5, seed=3); b = ctx.rand_normal((o,), 0, 0.5, seed=4)
outs = (ctx.empty((B, o)), ctx.empty((B, i)), ctx.empty((o, i)), ctx.empty((o,)))
for _ in range(steps): nn.fflayer_fwd_grad(X, W, b, dA, out=outs)
ctx.sync(); print("done")
