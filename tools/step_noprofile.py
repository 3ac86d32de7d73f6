#!/usr/bin/env python
"""Config-2 step timed with two CUDA events around 20 steps and NO per-kernel profiling events, next to the same loop with them."""
import os, sys, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tensor_ops_b200 as tb
from tensor_ops_b200 import nn
ctx = tb.Context(0)
B, i, o = 65536, 1024, 1024
X = ctx.rand_uniform((B, i), -1, 1, seed=1); dA = ctx.rand_normal((B, o), 0, 1, seed=2)
W = ctx.rand_normal((o, i), 0, 0.5, seed=3); b = ctx.rand_normal((o,), 0, 0.5, seed=4)
outs = (ctx.empty((B, o)), ctx.empty((B, i)), ctx.empty((o, i)), ctx.empty((o,)))
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
def run(n, prof):
    with torch.cuda.stream(stream):
        for _ in range(3): nn.fflayer_fwd_grad(X, W, b, dA, out=outs)
        torch.cuda.synchronize()
        if prof: ctx.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): nn.fflayer_fwd_grad(X, W, b, dA, out=outs)
        e1.record(); torch.cuda.synchronize()
        if prof: ctx.profile_summary(); ctx.profile(False)
        return e0.elapsed_time(e1) / n
a, bb = [], []
for _ in range(4):          # alternate, idling in between: some boxes power-cap after ~100 back-to-back steps
    time.sleep(1.5); a.append(run(20, False)); time.sleep(1.5); bb.append(run(20, True))
res = {"no_profile_ms": a, "with_profile_ms": bb}
print(json.dumps(res))
