#!/usr/bin/env python
"""HBM-bound kernel families of libtops_b200, each timed alone at a size far above L2 (CUDA events via tops_profile_*), reported as
algorithmic GB/s and as a fraction of the measured copy bandwidth (MEASURED_PEAKS.json hbm_gbs).  `--one NAME` runs a single family
(used under `ncu --set full` to capture dram__bytes.sum.per_second for profiles/).
   python tools/bench_bandwidth.py [--one NAME]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tensor_ops_b200 as tb
from tensor_ops_b200 import expr as E, _lib as L, nn
from tensor_ops_b200.tensor import CuTensor as T

ctx = tb.Context(0)
pk = os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")
PEAK = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
one = sys.argv[sys.argv.index("--one") + 1] if "--one" in sys.argv else None
import torch


def timed(fn, reps=10):
    for _ in range(3): fn()
    ctx.sync()
    stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn(); e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
    ctx.set_stream(None)
    return e0.elapsed_time(e1) / reps


N = 1 << 28                         # 1 GiB of fp32
R, Cc = 65536, 4096                 # [65536, 4096] fp32 = 1 GiB
cases = {}
def case(name, bytes_, build):
    cases[name] = (bytes_, build)

case("fill", 4 * N, lambda: (lambda t=ctx.empty((N,)): lambda: ctx.check(L.lib.tops_fill(ctx.h, t.b, 1.5)))())
case("axpy(scale)", 8 * N, lambda: (lambda x=ctx.rand_uniform((N,), -1, 1, seed=1), o=ctx.empty((N,)): lambda: T.scaleT(2.0, x))())
def _add():
    x = ctx.rand_uniform((N // 2,), -1, 1, seed=1); y = ctx.rand_uniform((N // 2,), -1, 1, seed=2)
    return lambda: T.sumT([x, y])
case("add (sumT of 2)", 12 * (N // 2), _add)
def _sgd():
    p = ctx.rand_uniform((N // 2,), -1, 1, seed=1); g = ctx.rand_uniform((N // 2,), -1, 1, seed=2)
    return lambda: nn.sgd_step([p], [g], 0.1)
case("sgd p - r*g", 12 * (N // 2), _sgd)
def _lift(f, nin):
    xs = [ctx.rand_uniform((N // 2,), 0.5, 1.5, seed=3 + j) for j in range(nin)]
    return lambda: T.liftT(f, xs)
case("lift logistic (catalogue)", 8 * (N // 2), lambda: _lift(E.logistic, 1))
case("lift d*logistic'(x) (catalogue)", 12 * (N // 2), lambda: _lift(lambda d, v: d * E.logistic_(v), 2))
case("lift generic (interpreter)", 12 * (N // 2), lambda: _lift(lambda a, b: E.tanh(a) * b + E.sqrt(b * b + 1), 2))
def _cast():
    x = ctx.rand_uniform((N // 2,), -1, 1, seed=1)
    return lambda: x.cast(L.BF16)
case("cast f32->bf16", 6 * (N // 2), _cast)
def _ger():
    x = ctx.rand_uniform((16384,), -1, 1, seed=1); y = ctx.rand_uniform((16384,), -1, 1, seed=2)
    return lambda: x._new(L.lib.tops_ger, x.b, y.b)
case("ger 16384x16384", 4 * 16384 * 16384, _ger)
def _bcast():
    row = ctx.rand_uniform((Cc,), -1, 1, seed=1)
    return lambda: T.broadcastRows(R, row)
case("broadcast_rows [65536,4096]", 4 * R * Cc, _bcast)
def _sumrows():
    x = ctx.rand_uniform((R, Cc), -1, 1, seed=1)
    return lambda: T.sumRows(x)
case("sum_rows [65536,4096]", 4 * R * Cc, _sumrows)
def _sum():
    x = ctx.rand_uniform((N,), -1, 1, seed=1)
    return lambda: x._new(L.lib.tops_sum, x.b)
case("sum (all)", 4 * N, _sum)
def _gemv(n, m):
    a = ctx.rand_uniform((n, m), -1, 1, seed=1); x = ctx.rand_uniform((m,), -1, 1, seed=2)
    return lambda: T.gmul(1, 1, 0, a, x)
case("gemv 16384x16384", 4 * 16384 * 16384, lambda: _gemv(16384, 16384))
case("gemv 1024x1024 (the reference's matVec)", 4 * 1024 * 1024, lambda: _gemv(1024, 1024))
def _transp3():
    x = ctx.rand_uniform((256, 1024, 1024), -1, 1, seed=1)
    return lambda: T.transp(x)
case("transp rank 3 [256,1024,1024] (tiled permute)", 8 * N, _transp3)
def _split():
    x = ctx.rand_uniform((65536, 1024), -1, 1, seed=1); d = ctx.rand_normal((65536, 1024), 0, 1, seed=2); w = ctx.rand_normal((1024, 1024), 0, 0.5, seed=3); b = ctx.rand_normal((1024,), 0, 0.5, seed=4)
    outs = (ctx.empty((65536, 1024)), ctx.empty((65536, 1024)), ctx.empty((1024, 1024)), ctx.empty((1024,)))
    def f():
        ctx.profile(True); nn.fflayer_fwd_grad(x, w, b, d, out=outs); p = ctx.profile_summary(); ctx.profile(False)
        f.ms = p["split_f16_X"]["ms"] / p["split_f16_X"]["launches"]
    return f
case("split_f16_rows(X) + max|dA| (inside fflayer_fwd_grad)", 3 * 4 * 65536 * 1024, _split)
def _softmax():
    z = ctx.rand_normal((1 << 22, 16), 0, 1, seed=1)
    return lambda: z._new(L.lib.tops_map_rows_softmax, z.b)
case("softmax rows [4Mi,16]", 8 * (1 << 26), _softmax)

for name, (nbytes, build) in cases.items():
    if one is not None and one not in name:
        continue
    fn = build()
    if "split_f16" in name:
        fn(); fn(); ms = fn.ms
    else:
        ms = timed(fn)
    print(json.dumps({"kernel": name, "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 3), "GB_s": round(nbytes / ms / 1e6, 1), "frac_of_measured_hbm_peak": round(nbytes / ms / 1e6 / PEAK, 3)}), flush=True)
    del fn
