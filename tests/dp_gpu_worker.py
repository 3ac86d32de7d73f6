"""Worker of tests/test_gpu_large.py::test_two_rank_data_parallel_paths_agree (one process per GPU under torch.distributed.run).
Every rank holds the same full batch, works on its own shard, and the three data-parallel schedules must all deliver the
oracle's FULL-batch [dW||db]:  (1) fwd_grad + NCCL all-reduce, (2) tops_fflayer_step_dp (all-reduce overlapped with dX),
(3) tops_fflayer_fwd_grad_mc (NVLS multimem.red fused into the dW GEMM; skipped if the platform has no multicast)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import tensor_ops_b200 as tb
from oracle import tensor_ops_oracle as O
from tensor_ops_b200 import dp, nn


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = tb.Context(local)
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    rng = np.random.default_rng(21)
    B, i, o = 2048 + 384, 512, 768             # ragged against the 256-row tiles and uneven across ranks
    X = rng.uniform(-1, 1, (B, i)).astype(np.float32); W = rng.normal(0, 0.5, (o, i)).astype(np.float32)
    b = rng.normal(0, 0.5, o).astype(np.float32); dA = rng.normal(size=(B, o)).astype(np.float32)
    ref = O.fflayer_logistic_dense(*(a.astype(np.float64) for a in (X, W, b, dA)))
    full = np.concatenate([ref[2].ravel(), ref[3]])
    lo, hi = dp.shard_range(B, rank, world)
    Xs, dAs = ctx.from_numpy(X[lo:hi]), ctx.from_numpy(dA[lo:hi])
    Wd, bd = ctx.from_numpy(W), ctx.from_numpy(b)
    layout = dp.PackedLayout.for_layers([(o, i)])
    rel = lambda g, r: float(np.linalg.norm(np.asarray(g, np.float64) - r) / np.linalg.norm(r))
    errs = {}
    # (1) plain NCCL after the GEMMs
    packed_t = torch.zeros(layout.numel, dtype=torch.float32, device=dev); packed = ctx.wrap_torch(packed_t)
    dWv, dbv = layout.views(packed)
    A, dX, _, _ = nn.fflayer_fwd_grad(Xs, Wd, bd, dAs, out=(None, None, dWv, dbv))
    dp.allreduce_sum_(packed_t); torch.cuda.synchronize()
    errs["nccl"] = rel(packed_t.cpu().numpy(), full)
    errs["A"] = rel(A.numpy(), ref[0][lo:hi]); errs["dX"] = rel(dX.numpy(), ref[1][lo:hi])
    # (2) overlapped schedule
    for r in (0, 8):
        ov = dp.OverlappedStep(ctx, layout, dev, reserve_sms=r)
        for _ in range(3):
            A2, dX2, g2 = ov.step(Xs, Wd, bd, dAs)
        torch.cuda.synchronize()
        errs[f"overlap{r}"] = rel(ov.packed_t.cpu().numpy(), full)
        errs[f"overlap{r}_dX"] = rel(dX2.numpy(), ref[1][lo:hi])
        ov.close()
    # (3) fused NVLS push
    try:
        fused = dp.FusedGradAllReduce(layout.numel, dev)
    except Exception as exc:
        fused = None
        if rank == 0:
            print("fused all-reduce unavailable:", exc)
    if fused is not None:
        local_grads = ctx.empty((layout.numel,))
        for _ in range(3):
            fused.begin()
            A3, dX3, _ = nn.fflayer_fwd_grad_mc(Xs, Wd, bd, dAs, fused.multicast_ptr, out=(None, None, local_grads))
            fused.end()
        torch.cuda.synchronize()
        errs["fused"] = rel(fused.local.cpu().numpy(), full)
        errs["fused_dX"] = rel(dX3.numpy(), ref[1][lo:hi])
    bad = {k: v for k, v in errs.items() if not v <= 1e-5}
    flag = torch.tensor([1.0 if not bad else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    print(f"rank {rank}: rows [{lo},{hi}) errs {errs}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if float(flag.item()) != 1.0:
        sys.exit(f"rank {rank}: errors above 1e-5: {bad}")
    if rank == 0:
        print("DP_GPU_WORKER_OK", flush=True)


if __name__ == "__main__":
    main()
