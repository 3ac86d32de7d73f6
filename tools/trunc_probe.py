#!/usr/bin/env python
"""Direction of the tensor core's accumulator truncation: signed mean relative error of an F16X3 product whose partial sums are all
positive, and of the same product with B negated (all partial sums negative), with long TMEM accumulation chunks so that the
truncation dominates.  toward zero: the two means have opposite signs (|result| too small both times); toward -inf: same sign."""
import os, sys, json
os.environ["TOPS_F16X3_CHUNK"] = "16"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tensor_ops_b200 as tb
ctx = tb.Context(0)
rng = np.random.default_rng(0)
M, K, N = 4096, 4096, 512
A = np.abs(rng.normal(size=(M, K))).astype(np.float32)
B = np.abs(rng.normal(size=(K, N))).astype(np.float32)
ref = A.astype(np.float64) @ B.astype(np.float64)
out = {}
for name, sgn in (("positive sums", 1.0), ("negative sums", -1.0)):
    got = ctx.from_numpy(A).gemm(ctx.from_numpy(sgn * B)).numpy().astype(np.float64)
    r = sgn * ref
    out[name] = {"mean_signed_rel_err": float(np.mean((got - r) / np.abs(r))), "mean_abs_rel_err": float(np.mean(np.abs(got - r) / np.abs(r)))}
print(json.dumps(out))
