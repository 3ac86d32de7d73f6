#!/usr/bin/env python
"""BASELINE.json configs[3]: ffLayer 4096->4096, GLOBAL batch 262144, bf16 storage / fp32 accumulate, batch-sharded over the GPUs
of one box (strong scaling: 262144/N rows per GPU) + one NCCL all-reduce of the fp32 [dW||db] buffer (64.02 MiB) per step.
   python tools/bench_cfg4_dp.py                      (1 GPU)
   python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_cfg4_dp.py
Prints one JSON line on rank 0 (not the driver's bench contract; numbers go to DESIGN.md / profiles/)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.stdout.flush(); real_stdout = os.fdopen(os.dup(1), "w"); os.dup2(2, 1)
import torch, torch.distributed as dist
import tensor_ops_b200 as tb
from tensor_ops_b200 import nn, dp, _lib as L

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
ctx = tb.Context(local)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
n, Bg = 4096, 262144
lo, hi = dp.shard_range(Bg, rank, world); B = hi - lo
def bf16(t): r = t.cast(L.BF16); return r
X = bf16(ctx.rand_uniform((B, n), -1, 1, seed=100 + rank)); dA = bf16(ctx.rand_normal((B, n), 0, 1, seed=200 + rank))
W = bf16(ctx.rand_normal((n, n), 0, 0.5 / 32, seed=1)); b = ctx.rand_normal((n,), 0, 0.5, seed=2)
A = ctx.empty((B, n), L.BF16); dX = ctx.empty((B, n), L.BF16)
layout = dp.PackedLayout.for_layers([(n, n)])
packed_t = torch.zeros(layout.numel, dtype=torch.float32, device=dev); packed = ctx.wrap_torch(packed_t)
dWv, dbv = layout.views(packed)
fused = None
local_grads = ctx.empty((layout.numel,))
if world > 1 and "--nccl" not in sys.argv:
    try: fused = dp.FusedGradAllReduce(layout.numel, dev)
    except Exception as exc: print("fused all-reduce unavailable:", exc, file=sys.stderr)
def step(comm=True):
    if comm and fused is not None:          # all-reduce fused into the dW / db epilogues (NVLS multimem.red)
        fused.begin(); nn.fflayer_fwd_grad_mc(X, W, b, dA, fused.multicast_ptr, out=(A, dX, local_grads)); fused.end()
        return
    nn.fflayer_fwd_grad(X, W, b, dA, out=(A, dX, dWv, dbv))
    if comm: dp.allreduce_sum_(packed_t)
def barrier():
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
def timed(comm, steps=10):
    for _ in range(3): step(comm)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): step(comm)
    e1.record(); barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    return ms
ms = timed(True); ms_nocomm = timed(False)
if rank == 0:
    peak = 1671.4
    try: peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception: pass
    flop = 6.0 * Bg * n * n
    real_stdout.write(json.dumps({"config": 4, "n_gpus": world, "scaling": "strong", "global_batch": Bg, "rows_per_gpu": B, "ms_per_step": ms,
                                  "ms_per_step_without_allreduce": ms_nocomm, "samples_per_s": Bg / ms * 1e3, "tflops_algorithmic_per_gpu": flop / world / ms / 1e9,
                                  "frac_of_measured_bf16_peak": flop / world / ms / 1e9 / peak, "allreduce_bytes": layout.numel * 4,
                                  "allreduce": "fused NVLS multimem.red in the GEMM epilogues" if fused is not None else "NCCL"}) + "\n")
    real_stdout.flush()
if world > 1: dist.destroy_process_group()
