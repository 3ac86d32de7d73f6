"""Data-parallel plumbing for the batched ffLayer/MLP gradient: one process per GPU, the batch axis is
sharded, parameters are replicated, and the parameter gradients of all layers travel in ONE packed fp32 buffer
`[dW0‖db0‖dW1‖db1‖…]` that is all-reduced (sum) once per step.

The reference has no parallelism of any kind (single process; `par`/`forkIO` appear nowhere) — the sum over samples
that `trainNetwork`'s fold performs one sample at a time (app/Dots.hs:74-80) is what the all-reduce completes here.
Only torch.distributed is used (NCCL on GPUs; gloo in the CPU tests); no device code lives in this module.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def shard_range(n_rows: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of a batch of `n_rows` samples owned by `rank`; the remainder goes to the
    first ranks, so sizes differ by at most one and every row is owned exactly once."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class PackedLayout:
    """Offsets of a list of parameter-shaped tensors inside one flat buffer."""

    def __init__(self, shapes: Sequence[Sequence[int]]):
        self.shapes = [tuple(int(d) for d in s) for s in shapes]
        self.offsets: List[int] = []
        off = 0
        for s in self.shapes:
            self.offsets.append(off)
            off += int(np.prod(s)) if len(s) else 1
        self.numel = off

    @staticmethod
    def for_layers(layer_dims: Sequence[Tuple[int, int]]) -> "PackedLayout":
        """layer_dims = [(o, i), ...]  ->  [dW0[o,i], db0[o], dW1, db1, ...] (the order netGrad returns them,
        FeedForward.hs:178-199)."""
        shapes: List[Tuple[int, ...]] = []
        for o, i in layer_dims:
            shapes += [(o, i), (o,)]
        return PackedLayout(shapes)

    def views(self, packed):
        """Split a CuTensor (via .view), a torch tensor or a NumPy array of `numel` elements into per-parameter views."""
        out = []
        for s, off in zip(self.shapes, self.offsets):
            n = int(np.prod(s)) if len(s) else 1
            if hasattr(packed, "view") and hasattr(packed, "ctx"):        # CuTensor
                out.append(packed.view(off, s))
            else:                                                          # torch / numpy: slicing keeps aliasing
                out.append(packed[off:off + n].reshape(s))
        return out


def allreduce_sum_(packed_torch, group=None):
    """In-place sum over ranks of the packed gradient buffer (torch tensor; NCCL on the current CUDA stream, or gloo on
    CPU).  No-op without an initialised process group (single GPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed_torch, op=dist.ReduceOp.SUM, group=group)
    return packed_torch


class FusedGradAllReduce:
    """The packed gradient buffer as NVLS symmetric memory: every rank allocates the same buffer, binds it to one multicast
    object (torch.distributed._symmetric_memory) and hands the MULTICAST address to `tops_fflayer_fwd_grad_mc`, whose GEMM
    epilogues emit `multimem.red` — the NVSwitch adds each split-K partial into every rank's replica while the GEMMs run, so the
    data-parallel all-reduce has no pass of its own.  Per step: `begin()` (zero + barrier), the fused call, `end()` (barrier).
    Raises RuntimeError when the platform has no multicast support (callers fall back to `allreduce_sum_`)."""

    def __init__(self, numel: int, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if not (dist.is_available() and dist.is_initialized()):
            raise RuntimeError("FusedGradAllReduce needs an initialised process group")
        group = group or dist.group.WORLD
        self.local = symm_mem.empty(numel, dtype=torch.float32, device=device)
        self.local.zero_()
        self.handle = symm_mem.rendezvous(self.local, group)
        self.multicast_ptr = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        if self.multicast_ptr == 0:
            raise RuntimeError("symmetric memory came up without a multicast pointer (no NVLS on this platform)")
        self.world = dist.get_world_size(group)

    def begin(self):
        """Zero the local replica and wait until every rank has done so (nobody's reductions may land in a stale buffer)."""
        self.local.zero_()
        self.handle.barrier(channel=0)

    def end(self):
        """All ranks' multimem reductions have been issued and completed: the local replica now holds the summed gradient."""
        self.handle.barrier(channel=1)


class OverlappedStep:
    """The data-parallel ffLayer step with the gradient all-reduce overlapped with the dX GEMM (tops_fflayer_step_dp).

    The library launches forward -> dW(+db) -> [event] -> dX on the compute stream; this helper owns a communication stream that
    waits for the event and runs the NCCL all-reduce of the packed [dW‖db] while dX is still computing (dX does not depend on
    dW).  `step()` returns after queueing everything; the compute stream is made to wait for the collective, so work queued
    after `step()` sees the summed gradient.  `reserve_sms` SMs are left to the collective by the persistent dX GEMM.
    Without a process group (single GPU) it degenerates to the plain call."""

    def __init__(self, ctx, layout: PackedLayout, device, reserve_sms: int = 8, group=None):
        import ctypes as C
        import torch
        from . import _lib as L
        self.ctx, self.layout, self.group, self.reserve_sms = ctx, layout, group, int(reserve_sms)
        self.packed_t = torch.zeros(layout.numel, dtype=torch.float32, device=device)
        self.packed = ctx.wrap_torch(self.packed_t)
        self.comm_stream = torch.cuda.Stream(device=device)
        ev = C.c_void_p()
        ctx.check(L.lib.tops_event_create(ctx.h, C.byref(ev)))
        self._ev = ev
        self._L, self._C, self._torch = L, C, torch

    def close(self):
        if self._ev is not None:
            self.ctx.check(self._L.lib.tops_event_destroy(self.ctx.h, self._ev))
            self._ev = None

    def step(self, X, W, b, dA, A=None, dX=None, act=None):
        """Queues one step; returns (A, dX, packed gradient CuTensor)."""
        import torch.distributed as dist
        L, C, torch = self._L, self._C, self._torch
        act = L.ACT_LOGISTIC if act is None else act
        from .tensor import CuTensor
        slots = [L.c_buf(A.b.value) if A is not None else L.c_buf(), L.c_buf(dX.b.value) if dX is not None else L.c_buf(), L.c_buf(self.packed.b.value)]
        multi = dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1
        self.ctx.check(L.lib.tops_fflayer_step_dp(self.ctx.h, X.b, W.b, b.b, act, dA.b, C.byref(slots[0]), C.byref(slots[1]), C.byref(slots[2]),
                                                  self._ev if multi else None, self.reserve_sms if multi else 0))
        if multi:
            main = torch.cuda.current_stream()
            self.ctx.check(L.lib.tops_stream_wait_event(self.ctx.h, C.c_void_p(self.comm_stream.cuda_stream), self._ev))
            with torch.cuda.stream(self.comm_stream):
                dist.all_reduce(self.packed_t, op=dist.ReduceOp.SUM, group=self.group)
            main.wait_stream(self.comm_stream)
        return (A if A is not None else CuTensor(self.ctx, slots[0]), dX if dX is not None else CuTensor(self.ctx, slots[1]), self.packed)
