#!/usr/bin/env python
"""Relative Frobenius error of the batched fwd+grad against the fp64 oracle for each precision mode (config-2 layer, 8192 rows)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tensor_ops_b200 as tb
from tensor_ops_b200 import nn
from oracle import tensor_ops_oracle as O
ctx = tb.Context(0)
rng = np.random.default_rng(0)
B, i, o = 8192, 1024, 1024
for init, wsd in (("reference init N(0,0.5^2)", 0.5), ("non-saturating init N(0,1/i)", 1.0 / np.sqrt(i))):
    X = rng.uniform(-1, 1, (B, i)).astype(np.float32); W = rng.normal(0, wsd, (o, i)).astype(np.float32)
    b = rng.normal(0, 0.5, o).astype(np.float32); dA = rng.normal(size=(B, o)).astype(np.float32)
    ref = O.fflayer_logistic_dense(*(a.astype(np.float64) for a in (X, W, b, dA)))
    for prec, name in ((tb.PREC_F16X3, "f16x3"), (tb.PREC_TF32_BF16X2, "tf32bf16"), (tb.PREC_TF32X3, "tf32x3"), (tb.PREC_TF32, "tf32"), (tb.PREC_FP32_SIMT, "simt")):
        ctx.set_precision(prec)
        got = nn.fflayer_fwd_grad(ctx.from_numpy(X), ctx.from_numpy(W), ctx.from_numpy(b), ctx.from_numpy(dA))
        errs = {n: float(np.linalg.norm(g.numpy().astype(np.float64) - r) / np.linalg.norm(r)) for n, g, r in zip(("A", "dX", "dW", "db"), got, ref)}
        print(json.dumps({"init": init, "precision": name, **{k: float(f"{v:.3e}") for k, v in errs.items()}}), flush=True)
