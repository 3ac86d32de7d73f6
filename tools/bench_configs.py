#!/usr/bin/env python
"""Measures the other BASELINE.json configs on one GPU (not the driver's bench line; numbers go into DESIGN.md / profiles/):
   config 3: MLP 784->512->256->10, logistic/logistic/softmax + crossEntropy, batch 32768, fp32 (netGrad over the batch)
   config 4: ffLayer 4096->4096, bf16 storage / fp32 accumulate, batch 262144 (per GPU share at 1 GPU), fwd+grad
   config 5: rank-3 contraction  x[64,64,64] . y[64,64] -> [64,64,64] -> sumRows -> [64,64], fwd + VJP
   usage: python tools/bench_configs.py [3] [4] [5]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import tensor_ops_b200 as tb
from tensor_ops_b200 import nn, _lib as L, top as TO
from tensor_ops_b200.tensor import CuTensor

ctx = tb.Context(0)
PEAK_BF16 = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"] if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 1590.0
PEAK_HBM = 6546.2


def timeit(fn, steps=10, warm=3):
    """(ms per step over `steps` back-to-back calls with one synchronise at the end, per-kernel device ms from a second, profiled pass —
    the profiling events sit between the launches and would otherwise be part of the step time)"""
    for _ in range(warm): fn()
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(steps): fn()
    ctx.sync()
    wall = (time.perf_counter() - t0) / steps * 1e3
    ctx.profile(True)
    for _ in range(steps): fn()
    prof = ctx.profile_summary(); ctx.profile(False)     # synchronises
    return wall, {k: v["ms"] / v["launches"] * (v["launches"] / steps) for k, v in prof.items()}


only = None
if "--prec" in sys.argv:            # e.g. --prec f16x3: one precision only (used under ncu)
    k = sys.argv.index("--prec"); only = sys.argv[k + 1]; del sys.argv[k:k + 2]
PRECS_ALL = ((tb.PREC_F16X3, "f16x3"), (tb.PREC_TF32_BF16X2, "tf32bf16"), (tb.PREC_TF32X3, "tf32x3"), (tb.PREC_TF32, "tf32"), (tb.PREC_FP32_SIMT, "simt"))
def precs(names): return [(p, n) for p, n in PRECS_ALL if n in names and (only is None or n == only)]
which = set(sys.argv[1:]) or {"3", "4", "5"}
if "3" in which:
    B, dims = 32768, [784, 512, 256, 10]
    Ws = [ctx.rand_normal((dims[l + 1], dims[l]), 0, 0.5 / np.sqrt(dims[l]) * 2, seed=10 + l) for l in range(3)]
    bs = [ctx.rand_normal((dims[l + 1],), 0, 0.5, seed=20 + l) for l in range(3)]
    X = ctx.rand_uniform((B, 784), 0, 1, seed=1)
    Yh = np.zeros((B, 10), np.float32); Yh[np.arange(B), np.random.default_rng(0).integers(0, 10, B)] = 1
    Y = ctx.from_numpy(Yh)
    acts = [L.ACT_LOGISTIC, L.ACT_LOGISTIC, L.ACT_SOFTMAX]
    for prec, name in precs(("f16x3", "tf32bf16", "tf32x3", "tf32")):
        ctx.set_precision(prec)
        ms, per = timeit(lambda: nn.mlp_fwd_grad(Ws, bs, acts, L.LOSS_CROSS_ENTROPY, X, Y), steps=30)
        flop = 6.0 * B * (784 * 512 + 512 * 256 + 256 * 10)
        with ctx.record(arena_bytes=3 << 30) as g:     # the whole netGrad recorded once, replayed as one graph launch
            nn.mlp_fwd_grad(Ws, bs, acts, L.LOSS_CROSS_ENTROPY, X, Y)
        for _ in range(5): g.launch()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(50): g.launch()
        ctx.sync()
        gms = (time.perf_counter() - t0) / 50 * 1e3
        print(json.dumps({"config": 3, "precision": name, "ms_per_step": ms, "samples_per_s": B / ms * 1e3, "tflops_algorithmic": flop / ms / 1e9,
                          "frac_of_tf32_peak": flop / ms / 1e9 / (PEAK_BF16 / 2),
                          "graph_replay": {"ms_per_step": gms, "kernels": g.kernel_count(), "samples_per_s": B / gms * 1e3, "frac_of_tf32_peak": flop / gms / 1e9 / (PEAK_BF16 / 2)},
                          "per_step_kernel_ms": per}), flush=True)
        g.close()
if "4" in which:
    B, n = 262144, 4096
    Xf = ctx.rand_uniform((B, n), -1, 1, seed=1); X = Xf.cast(L.BF16); del Xf
    dAf = ctx.rand_normal((B, n), 0, 1, seed=2); dA = dAf.cast(L.BF16); del dAf
    Wf = ctx.rand_normal((n, n), 0, 0.5 / 64 * 2, seed=3); W = Wf.cast(L.BF16); del Wf
    b = ctx.rand_normal((n,), 0, 0.5, seed=4)
    outs = (ctx.empty((B, n), L.BF16), ctx.empty((B, n), L.BF16), ctx.empty((n, n)), ctx.empty((n,)))
    ms, per = timeit(lambda: nn.fflayer_fwd_grad(X, W, b, dA, out=outs), steps=5)
    flop = 6.0 * B * n * n
    print(json.dumps({"config": 4, "precision": "bf16 storage, fp32 accumulate", "ms_per_step": ms, "samples_per_s": B / ms * 1e3,
                      "tflops_algorithmic": flop / ms / 1e9, "frac_of_bf16_peak": flop / ms / 1e9 / PEAK_BF16, "per_step_kernel_ms": per}), flush=True)
if "5" in which:
    rng = np.random.default_rng(5)
    x = ctx.from_numpy(rng.normal(size=(64, 64, 64))); y = ctx.from_numpy(rng.normal(size=(64, 64))); d = ctx.from_numpy(rng.normal(size=(64, 64)))
    op = TO.compose(TO.sumRows(), TO.inner(2, 1))      # inner (LS (LS LZ)) (LS LZ) >>> sumRows   (SURVEY §8-d note on config 5)
    for prec, name in precs(("f16x3", "tf32bf16", "tf32x3", "simt")):
        ctx.set_precision(prec)
        def step():
            TO.runTOp(op, [x, y]); TO.gradTOp_(op, [x, y], [d])
        n0 = ctx.launch_count(); step(); launches = ctx.launch_count() - n0
        ms, per = timeit(step, steps=20)
        bytes_alg = 3.1 * 2 ** 20
        # the same step recorded once and replayed as one CUDA graph launch (tops_graph_*): device time per replay
        with ctx.record() as g:
            step()
        for _ in range(20): g.launch()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(500): g.launch()
        ctx.sync()
        gms = (time.perf_counter() - t0) / 500 * 1e3
        # a graph launch itself costs ~12 us of launch pipeline, more than the two kernels: 32 steps recorded into ONE graph show the device time
        with ctx.record() as g32:
            for _ in range(32): step()
        for _ in range(5): g32.launch()
        ctx.sync()
        t0 = time.perf_counter()
        for _ in range(100): g32.launch()
        ctx.sync()
        g32ms = (time.perf_counter() - t0) / 100 / 32 * 1e3
        g32k = g32.kernel_count(); g32.close()
        print(json.dumps({"config": 5, "precision": name, "ms_per_step": ms, "kernel_launches_per_step": launches, "algorithmic_GB_s": bytes_alg / ms / 1e6,
                          "frac_of_hbm_peak": bytes_alg / ms / 1e6 / PEAK_HBM,
                          "graph_replay": {"ms_per_step": gms, "kernels": g.kernel_count(), "algorithmic_GB_s": bytes_alg / gms / 1e6, "frac_of_hbm_peak": bytes_alg / gms / 1e6 / PEAK_HBM},
                          "graph_replay_32_steps_per_graph": {"ms_per_step": g32ms, "kernels": g32k, "algorithmic_GB_s": bytes_alg / g32ms / 1e6, "frac_of_hbm_peak": bytes_alg / g32ms / 1e6 / PEAK_HBM},
                          "note": "3.1 MiB of algorithmic traffic per step; eager = one API call per Tensor method from Python, graph_replay = the recorded step, one graph launch per step (bound by the launch itself); graph_replay_32_steps_per_graph = 32 steps recorded into one graph", "per_step_kernel_ms": per}), flush=True)
        g.close()
