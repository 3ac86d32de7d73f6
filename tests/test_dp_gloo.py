"""World-size-2 (and 3) CPU test of the data-parallel plumbing (tensor_ops_b200.dp) over gloo: each rank evaluates the
ORACLE's batched gradient on its own batch shard, packs [dW‖db] exactly as bench.py does, all-reduces, and must recover
the full-batch gradient.  The device kernels are covered by tests/test_gpu_parity.py; this covers the N>1 host logic."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import dp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, B, dims, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)             # same parameters and the same FULL batch on every rank
        i0 = dims[0][1]
        Ws = [rng.normal(0, 0.5, d) for d in dims]; bs = [rng.normal(0, 0.5, d[0]) for d in dims]
        X = rng.uniform(-1, 1, (B, i0)); Y = rng.uniform(0, 1, (B, dims[-1][0]))
        acts = ["logistic"] * len(dims)
        lo, hi = dp.shard_range(B, rank, world)
        lay = dp.PackedLayout.for_layers(dims)
        packed = torch.zeros(lay.numel, dtype=torch.float64)
        views = lay.views(packed)
        if hi > lo:
            _, _, _, dWs, dbs = O.mlp_dense_fwd_grad(X[lo:hi], Ws, bs, acts, "squaredError", Y[lo:hi])
            for l in range(len(dims)):
                views[2 * l].copy_(torch.from_numpy(dWs[l])); views[2 * l + 1].copy_(torch.from_numpy(dbs[l]))
        dp.allreduce_sum_(packed)
        _, _, _, dWf, dbf = O.mlp_dense_fwd_grad(X, Ws, bs, acts, "squaredError", Y)
        full = np.concatenate([np.concatenate([dWf[l].ravel(), dbf[l].ravel()]) for l in range(len(dims))])
        err = float(np.linalg.norm(packed.numpy() - full) / np.linalg.norm(full))
        q.put((rank, lo, hi, err))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 64), (2, 7), (3, 10)])
def test_sharded_gradient_allreduce_matches_full_batch(world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    dims = [(6, 5), (3, 6)]
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, dims, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rows = sorted((lo, hi) for _, lo, hi, _ in res)
    assert rows[0][0] == 0 and rows[-1][1] == B and all(rows[k][1] == rows[k + 1][0] for k in range(world - 1))
    assert all(err < 1e-12 for *_, err in res), res


def test_shard_range_partitions_every_row_once():
    for n in (0, 1, 7, 64, 65537):
        for w in (1, 2, 3, 8):
            r = [dp.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_packed_layout_offsets():
    lay = dp.PackedLayout.for_layers([(4, 3), (2, 4)])
    assert lay.shapes == [(4, 3), (4,), (2, 4), (2,)] and lay.offsets == [0, 12, 16, 24] and lay.numel == 26
    a = np.arange(26.0)
    v = lay.views(a)
    v[2][...] = -1
    assert (a[16:24] == -1).all()          # views alias the packed buffer


def test_fused_allreduce_needs_a_process_group():
    """dp.FusedGradAllReduce (NVLS symmetric memory) must fail loudly, not silently degrade, outside a distributed job."""
    with pytest.raises(RuntimeError):
        dp.FusedGradAllReduce(16, torch.device("cpu"))
