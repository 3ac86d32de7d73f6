#!/usr/bin/env python
"""Host-side enqueue time of one netGrad call (config 3) versus its device time: is the eager path launch-bound?"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tensor_ops_b200 as tb
from tensor_ops_b200 import nn, _lib as L
ctx = tb.Context(0)
B, dims = 32768, [784, 512, 256, 10]
Ws = [ctx.rand_normal((dims[l + 1], dims[l]), 0, 1 / np.sqrt(dims[l]), seed=10 + l) for l in range(3)]
bs = [ctx.rand_normal((dims[l + 1],), 0, 0.5, seed=20 + l) for l in range(3)]
X = ctx.rand_uniform((B, 784), 0, 1, seed=1)
Yh = np.zeros((B, 10), np.float32); Yh[np.arange(B), np.random.default_rng(0).integers(0, 10, B)] = 1
Y = ctx.from_numpy(Yh)
acts = [L.ACT_LOGISTIC, L.ACT_LOGISTIC, L.ACT_SOFTMAX]
f = lambda: nn.mlp_fwd_grad(Ws, bs, acts, L.LOSS_CROSS_ENTROPY, X, Y)
for _ in range(5): f()
ctx.sync()
n = 50
t0 = time.perf_counter()
for _ in range(n): r = f()
t1 = time.perf_counter()
ctx.sync()
t2 = time.perf_counter()
n0 = ctx.launch_count(); f(); launches = ctx.launch_count() - n0
print(json.dumps({"enqueue_ms_per_call": (t1 - t0) / n * 1e3, "total_ms_per_call": (t2 - t0) / n * 1e3, "kernel_launches": launches}))
# ---- where the host time goes: the bare C call, and the unit costs of a launch / an allocation through the same library
import ctypes as C
from tensor_ops_b200.nn import _arr
from tensor_ops_b200.tensor import CuTensor
def bare():
    nl = 3
    A, ls, dX = L.c_buf(), L.c_buf(), L.c_buf()
    dW = (L.c_buf * nl)(); db = (L.c_buf * nl)(); a = (C.c_int * nl)(*acts)
    t = time.perf_counter()
    rc = L.lib.tops_mlp_fwd_grad(ctx.h, nl, _arr(Ws), _arr(bs), a, L.LOSS_CROSS_ENTROPY, X.b, Y.b, C.byref(A), C.byref(ls), C.byref(dX), dW, db)
    dt = time.perf_counter() - t
    outs = [CuTensor(ctx, A), CuTensor(ctx, ls), CuTensor(ctx, dX)] + [CuTensor(ctx, L.c_buf(dW[l])) for l in range(nl)] + [CuTensor(ctx, L.c_buf(db[l])) for l in range(nl)]
    return dt, outs
ctx.sync()
ts = []
for _ in range(30):
    dt, outs = bare(); ts.append(dt); ctx.sync()
small = ctx.empty((1024,))
ctx.sync(); t0 = time.perf_counter()
for _ in range(2000): ctx.check(L.lib.tops_fill(ctx.h, small.b, 1.0))
t_fill = (time.perf_counter() - t0) / 2000; ctx.sync()
t0 = time.perf_counter()
for _ in range(2000): ctx.empty((1024,))
t_alloc = (time.perf_counter() - t0) / 2000; ctx.sync()
print(json.dumps({"bare_c_call_ms_median": float(np.median(ts)) * 1e3, "bare_c_call_ms_min": float(np.min(ts)) * 1e3,
                  "fill_launch_us_incl_ctypes": t_fill * 1e6, "alloc_free_us_incl_python": t_alloc * 1e6}))
