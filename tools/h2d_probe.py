#!/usr/bin/env python
"""Pinned host -> device copy bandwidth on this box: one stream vs two concurrent streams, and chunk-size sensitivity."""
import json, time, torch
dev = torch.device("cuda:0")
n = 512 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
def bw(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return n * reps / (time.perf_counter() - t0) / 1e9
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one(): d.copy_(h, non_blocking=True)
def two():
    with torch.cuda.stream(s1): d[: n // 2].copy_(h[: n // 2], non_blocking=True)
    with torch.cuda.stream(s2): d[n // 2 :].copy_(h[n // 2 :], non_blocking=True)
def chunks(k):
    def f():
        c = n // k
        for i in range(k): d[i * c:(i + 1) * c].copy_(h[i * c:(i + 1) * c], non_blocking=True)
    return f
print(json.dumps({"one_stream_GB_s": bw(one), "two_streams_GB_s": bw(two), "16_chunks_GB_s": bw(chunks(16)), "64_chunks_GB_s": bw(chunks(64))}))
