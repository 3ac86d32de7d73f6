"""H2D bandwidth: torch pinned vs tops_host_alloc (default / write-combined), 256 MiB copies through tops_upload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import tensor_ops_b200 as tb
ctx = tb.Context(0)
B, i = 65536, 1024
dev = ctx.empty((B, i))
src = np.random.default_rng(0).standard_normal((B, i)).astype(np.float32)
bufs = {"torch pinned": torch.empty((B, i), dtype=torch.float32, pin_memory=True).numpy(),
        "tops_host_alloc": ctx.host_empty((B, i), False), "tops_host_alloc write-combined": ctx.host_empty((B, i), True)}
for name, h in bufs.items():
    h[...] = src
    for _ in range(2): dev.upload(h)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(10): dev.upload(h)
    ctx.sync()
    dt = (time.perf_counter() - t0) / 10
    print(f"{name}: {h.nbytes / dt / 1e9:.1f} GB/s", flush=True)
