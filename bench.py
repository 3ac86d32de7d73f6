#!/usr/bin/env python
"""bench.py — ffLayer forward+gradient throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision f16x3|tf32bf16|tf32x3|tf32|simt]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one batched runTOp + gradTOp' of `ffLayer' >>> logistic`: X[B,i], W[o,i], b[o], dA[B,o] ->
A[B,o], dX[B,i], dW[o,i], db[o]; with N > 1 every rank processes its own batch shard of B rows (weak scaling) and the
packed [dW‖db] buffer is all-reduced once per step over NCCL.  Workload = BASELINE.json configs[1]:
i = o = 1024, B = 65536 per GPU, fp32 storage.

  value     samples/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the host-buffer entry point: every step copies X and dA from pinned host memory to
            HBM and reads the gradient [dW‖db] back to the host
  roofline  the dominant kernel (the largest of the three tcgen05 GEMMs) timed with CUDA events on its own stream
            inside the timed region (tops_profile_*), algorithmic FLOP = 2*B*i*o per GEMM
  parity    relative Frobenius error of A / dX / dW / db against the fp64 oracle on a 4096-row slice of the TIMED inputs in
            the TIMED precision mode; the line is refused (exit 3) when any exceeds 1e-5
  cpu_baseline / --impl reference   the oracle's per-sample restatement of the hmatrix op sequence (the reference is
            Haskell and cannot be built here), on the host cores, on a bounded sample of the same workload; `variants` adds
            the 1-thread fp64 run (what the apps instantiate) and a best-effort batched sgemm formulation on all cores
  extra.config4   BASELINE configs[3] (4096->4096, bf16, GLOBAL batch 262144 sharded over the N GPUs: strong scaling)
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CFG = {"i": 1024, "o": 1024, "B": 65536}
WORKLOAD = "ffLayer 1024->1024 logistic, batch 65536 per GPU, fp32, runTOp+gradTOp' (BASELINE configs[1])"
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            flat = {}

            def walk(x):
                if isinstance(x, dict):
                    for k, v in x.items():
                        if isinstance(v, (int, float)):
                            flat.setdefault(k, float(v))
                        else:
                            walk(v)
            walk(d)
            if "bf16_tflops" in flat:
                return flat, "measured"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback"


NCU_SUMMARY = {"f16x3": ("profiles/r2_gemm_f16x3_ncu.csv", "<__half, "), "tf32bf16": ("profiles/r1_gemm_parity_ncu.csv", "<float, ")}


def ncu_traffic_bytes(tag, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel.  ncu cannot run inside a timed bench, so
    this is STATIC: read from the committed `ncu --set full` summary of the same bench command (tools/run_ncu_gemm.sh); returns
    (bytes, file) or (None, None) when no summary exists for this precision mode."""
    rel, prefix = NCU_SUMMARY.get(precision, (None, None))
    if rel is None:
        return None, None
    path = os.path.join(ROOT, rel)
    want = {"gemm_fwd": prefix + "0, 0,", "gemm_dW": prefix + "1, 1,", "gemm_dX": prefix + "0, 1,"}.get(tag)
    if not want or not os.path.exists(path):
        return None, None
    try:
        import csv
        rows = {r[0]: r for r in csv.reader(open(path))}
        names = rows["Kernel Name"][2:]
        col = next(k for k, n in enumerate(names) if want in n)
        rd, wr = rows["dram__bytes_read.sum"], rows["dram__bytes_write.sum"]
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
        return float(rd[2 + col]) * scale.get(rd[1], 1.0) + float(wr[2 + col]) * scale.get(wr[1], 1.0), rel
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML, 50 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.active = True   # cleared while the host prepares buffers between timed regions (idle clocks are not load)
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80,
                 "sw_power_cap": 0x4, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self._stop_evt.is_set():
            if not self.active:
                self._stop_evt.wait(0.002)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.004)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------- CPU arms
def cpu_step_all_cores(O, X, W, b, dA, threads):
    """The reference's per-sample op sequence (oracle.cpu_fflayer_step_reference: gemv, axpy, cmap logistic, recomputed forward,
    ger, gemv (tr W)) over a batch.  The reference itself is single-threaded (no par/forkIO anywhere); to give the CPU arm every
    host core, samples are split across `threads` Python threads (NumPy/OpenBLAS release the GIL; BLAS pinned to 1 thread per
    call to avoid oversubscription) and the per-thread parameter gradients are summed."""
    if threads <= 1:
        return O.cpu_fflayer_step_reference(X, W, b, dA)
    from concurrent.futures import ThreadPoolExecutor
    import numpy as np
    bounds = np.linspace(0, X.shape[0], threads + 1).astype(int)
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(lambda k: O.cpu_fflayer_step_reference(X[bounds[k]:bounds[k + 1]], W, b, dA[bounds[k]:bounds[k + 1]]), range(threads)))
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), sum(p[2] for p in parts), sum(p[3] for p in parts))


def _blas_single_thread():
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def cpu_reference_rate(seconds_target, i, o, dtype_name="float32", max_samples=16384):
    """Times the CPU port on a bounded sample of the workload with all host cores; returns (samples/s, n_samples, seconds, threads)."""
    import numpy as np
    from oracle import tensor_ops_oracle as O
    threads = os.cpu_count() or 1
    dt = np.dtype(dtype_name)
    rng = np.random.default_rng(0)
    W = rng.normal(0, 0.5, (o, i)).astype(dt); b = rng.normal(0, 0.5, o).astype(dt)
    with _blas_single_thread():
        probe = 16 * threads
        X = rng.uniform(-1, 1, (probe, i)).astype(dt); dA = rng.standard_normal((probe, o)).astype(dt)
        cpu_step_all_cores(O, X, W, b, dA, threads)
        t0 = time.perf_counter(); cpu_step_all_cores(O, X, W, b, dA, threads); per = (time.perf_counter() - t0) / probe
        n = int(max(32, min(max_samples, seconds_target / max(per, 1e-9))))
        X = rng.uniform(-1, 1, (n, i)).astype(dt); dA = rng.standard_normal((n, o)).astype(dt)
        t0 = time.perf_counter(); cpu_step_all_cores(O, X, W, b, dA, threads); dtm = time.perf_counter() - t0
    return n / dtm, n, dtm, threads


def cpu_baseline_variants(i, o):
    """SURVEY 8-d's two other CPU numbers: (1) the same per-sample op sequence in fp64 on ONE thread — what the reference's
    apps run (`HMat Double`, app/Dots.hs:145; the code has no parallel construct); (2) a best-effort CPU formulation the
    reference does not have: the three batched sgemm calls Z = X W^T, dW = dZ^T X, dX = dZ W on all cores."""
    import numpy as np
    from oracle import tensor_ops_oracle as O
    out = {}
    rng = np.random.default_rng(1)
    W = rng.normal(0, 0.5, (o, i)); b = rng.normal(0, 0.5, o)
    with _blas_single_thread():
        n = 96
        X = rng.uniform(-1, 1, (n, i)); dA = rng.standard_normal((n, o))
        O.cpu_fflayer_step_reference(X[:8], W, b, dA[:8])
        t0 = time.perf_counter(); O.cpu_fflayer_step_reference(X, W, b, dA); dt = time.perf_counter() - t0
        out["port_fp64_1thread"] = {"value": n / dt, "unit": "samples/s", "cores": 1, "kind": "port",
                                    "sample": f"{n} samples, per-sample hmatrix op sequence in fp64 (the dtype the reference's apps instantiate), 1 thread"}
    n = 8192
    X = rng.uniform(-1, 1, (n, i)).astype(np.float32); dA = rng.standard_normal((n, o)).astype(np.float32)
    Wf, bf = W.astype(np.float32), b.astype(np.float32)

    def dense():
        A = 1.0 / (1.0 + np.exp(-(X @ Wf.T + bf)))
        dZ = dA * (A * (1.0 - A))
        return A, dZ @ Wf, dZ.T @ X, dZ.sum(axis=0)
    dense()
    t0 = time.perf_counter(); dense(); dt = time.perf_counter() - t0
    out["batched_sgemm_allcores"] = {"value": n / dt, "unit": "samples/s", "cores": os.cpu_count() or 1, "kind": "port",
                                     "sample": f"{n} samples, dense closed form (3 sgemm + elementwise) in NumPy/OpenBLAS fp32 on all cores — not how the reference computes, a best-effort CPU number"}
    return out


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference is Haskell (no GHC in this
    image, no C sources to compile), so this executes the oracle port of its hmatrix op sequence on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import tensor_ops_oracle as O
    i, o = CFG["i"], CFG["o"]
    threads = os.cpu_count() or 1
    rng = np.random.default_rng(0)
    W = rng.normal(0, 0.5, (o, i)).astype(np.float32); b = rng.normal(0, 0.5, o).astype(np.float32)
    n = 64 * threads   # samples per step: a bounded sample of the 65536-row batch
    X = rng.uniform(-1, 1, (n, i)).astype(np.float32); dA = rng.standard_normal((n, o)).astype(np.float32)
    with _blas_single_thread():
        for _ in range(args.warmup):
            cpu_step_all_cores(O, X[:4 * threads], W, b, dA[:4 * threads], threads)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_step_all_cores(O, X, W, b, dA, threads)
        dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {"impl": "reference", "metric": "ffLayer fwd+grad samples/sec", "value": v, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": f"{n} samples per step of the 65536-sample batch"},
            "cpu_baseline": {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                             "sample": f"{n} samples/step x {args.steps} steps, per-sample hmatrix op sequence (oracle.cpu_fflayer_step_reference), NumPy/OpenBLAS fp32, samples split over {threads} threads"},
            "e2e": {"value": v, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
PREC_TEXT = {"f16x3": "every fp32 operand stored once as an fp16 pair (hi, lo) of x*2^k; three fp16 tcgen05 passes hi*hi + lo*hi + hi*lo, chunked fp32 accumulation (fp32-grade: parity mode)",
             "tf32bf16": "TF32 hi*hi + two bf16 correction passes on tcgen05 (fp32-grade ~1.4e-6, parity mode)",
             "tf32x3": "3xTF32 split on tcgen05 (fp32-grade, parity mode)", "tf32": "single-pass TF32 on tcgen05 (throughput mode, ~7e-4 rel err)", "simt": "fp32 FFMA"}
# tensor-core work per algorithmic FLOP in TF32-pass equivalents (fp16/bf16 MMAs run at twice the TF32 rate)
PASSES = {"f16x3": 1.5, "tf32bf16": 2.0, "tf32x3": 3.0, "tf32": 1.0, "simt": 0.0}


def bind_to_gpu_numa_node(local):
    """e2e path: pin this rank (and therefore the first-touch placement of its pinned staging buffers) to the CPUs of the GPU's
    NUMA node, so that 8 ranks do not pull their 512 MiB/step through one memory controller.  Returns a description or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local)
        n_words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {w * 64 + bit for w, word in enumerate(mask) for bit in range(64) if (word >> bit) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} CPUs of GPU {local}'s NUMA node (NVML cpu affinity)"
        return f"no narrower affinity available ({len(allowed)} CPUs allowed, GPU-local set has {len(cpus)})"
    except Exception as exc:
        return f"unavailable ({type(exc).__name__})"


def measure_tf32_peak(torch, dev):
    """cuBLAS TF32 GEMM, 8192^3, best of 10 — the same way the driver measures the bf16 figure in MEASURED_PEAKS.json (which has no
    TF32 entry).  Runs AFTER the timed regions."""
    n = 8192
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(n, n, device=dev); b = torch.randn(n, n, device=dev)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); a @ b; e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2.0 * n ** 3 / best / 1e9
    except Exception:
        return None
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def parity_check(ctx, nn, np, X, W, b, dA, A, dX, rows=4096):
    """Relative Frobenius error vs the fp64 oracle on the first `rows` rows of the TIMED inputs, in the TIMED precision mode:
    A and dX rows are taken from the timed full-batch result; dW and db come from one more call on the row slice (they are sums
    over the batch, so the slice has its own)."""
    from oracle import tensor_ops_oracle as O
    rows = min(rows, X.shape[0])
    i, o = X.shape[1], W.shape[0]
    Xs, dAs = X.view(0, (rows, i)), dA.view(0, (rows, o))
    Xh, dAh, Wh, bh = (t.numpy().astype(np.float64) for t in (Xs, dAs, W, b))
    ref = O.fflayer_logistic_dense(Xh, Wh, bh, dAh)
    got = nn.fflayer_fwd_grad(Xs, W, b, dAs)
    rel = lambda g, r: float(np.linalg.norm(np.asarray(g, dtype=np.float64) - r) / max(np.linalg.norm(r), 1e-300))
    out = {"rows": rows,
           "A": rel(A.view(0, (rows, o)).numpy(), ref[0]), "dX": rel(dX.view(0, (rows, i)).numpy(), ref[1]),
           "dW": rel(got[2].numpy(), ref[2]), "db": rel(got[3].numpy(), ref[3]),
           "A_slice_call": rel(got[0].numpy(), ref[0]), "dX_slice_call": rel(got[1].numpy(), ref[1]),
           "tolerance": 1e-5, "metric": "||dev - oracle_fp64||_F / ||oracle_fp64||_F"}
    out["ok"] = all(out[k] <= 1e-5 for k in ("A", "dX", "dW", "db", "A_slice_call", "dX_slice_call"))
    return out


def config4_extra(ctx, tb, nn, dp, torch, dist, dev, rank, world, peaks):
    """BASELINE configs[3]: ffLayer 4096->4096, GLOBAL batch 262144 sharded over the ranks (strong scaling), bf16 storage with fp32
    accumulation, fp32 [dW‖db] (64.02 MiB) all-reduced once per step, overlapped with the dX GEMM."""
    from tensor_ops_b200 import _lib as L
    n, Bg = 4096, 262144
    lo, hi = dp.shard_range(Bg, rank, world); B = hi - lo
    X = ctx.rand_uniform((B, n), -1, 1, seed=300 + rank).cast(L.BF16); dA = ctx.rand_normal((B, n), 0, 1, seed=400 + rank).cast(L.BF16)
    W = ctx.rand_normal((n, n), 0, 0.5, seed=3).cast(L.BF16); b = ctx.rand_normal((n,), 0, 0.5, seed=4)
    A = ctx.empty((B, n), L.BF16); dX = ctx.empty((B, n), L.BF16)
    layout = dp.PackedLayout.for_layers([(n, n)])
    ov = dp.OverlappedStep(ctx, layout, dev, reserve_sms=8)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    steps = 5
    for _ in range(3):
        ov.step(X, W, b, dA, A=A, dX=dX)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ov.step(X, W, b, dA, A=A, dX=dX)
    e1.record(); barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ov.close()
    flop = 6.0 * Bg * n * n
    return {"workload": "BASELINE configs[3]: ffLayer 4096->4096 logistic, GLOBAL batch 262144, bf16 storage / fp32 accumulate, batch-sharded, [dW||db] fp32 all-reduce overlapped with dX",
            "scaling": "strong", "n_gpus": world, "rows_per_gpu": B, "ms_per_step": ms, "value": Bg / (ms * 1e-3), "unit": "samples/s", "steps": steps, "warmup": 3,
            "tflops_per_gpu": flop / world / (ms * 1e-3) / 1e12, "frac_of_measured_bf16_peak": flop / world / (ms * 1e-3) / 1e12 / peaks["bf16_tflops"],
            "allreduce_bytes": layout.numel * 4, "init": "W, b ~ N(0, 0.5^2) (FeedForward.hs:206-207), X ~ U(-1,1), dA ~ N(0,1), rounded to bf16"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="f16x3", choices=["f16x3", "tf32bf16", "tf32x3", "tf32", "simt"])
    ap.add_argument("--batch", type=int, default=CFG["B"], help="rows per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-side", action="store_true", help="skip the single-pass TF32 side measurement")
    ap.add_argument("--no-extra", action="store_true", help="skip the config-4 sub-record and the TF32 peak measurement")
    ap.add_argument("--allreduce", default="auto", choices=["auto", "nccl", "overlap", "fused"],
                    help="N>1: auto = time the candidates (NCCL after the GEMMs / NCCL overlapped with dX / fused NVLS push) and keep the fastest")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference_arm(args)

    # stdout carries exactly ONE line (the JSON, rank 0): libraries that print to fd 1 (NCCL's "NCCL version ..." banner) are
    # sent to stderr for the rest of the run, and the JSON line is written to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa_node(local) if not args.no_e2e else None   # before the first allocation: pinned pages are first-touch

    import numpy as np
    import torch
    import torch.distributed as dist
    import tensor_ops_b200 as tb
    from tensor_ops_b200 import nn, dp

    if args.gpus != world:
        if world == 1 and args.gpus > 1:
            sys.exit("bench.py --gpus N>1 must be launched with torch.distributed.run (one process per GPU)")
    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device — tensor_ops_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    ctx = tb.Context(local)
    # one explicit (non-default) torch stream carries everything: the library's kernels, the NCCL all-reduce and the
    # timing events.  (The legacy default stream has handle 0, which tops_set_stream reads as "use the context's own".)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    prec = {"f16x3": tb.PREC_F16X3, "tf32bf16": tb.PREC_TF32_BF16X2, "tf32x3": tb.PREC_TF32X3, "tf32": tb.PREC_TF32, "simt": tb.PREC_FP32_SIMT}[args.precision]
    ctx.set_precision(prec)

    B, i, o = args.batch, CFG["i"], CFG["o"]
    # synthetic inputs at the reference's distributions (FeedForward.hs:206-207, Dots.hs:63), generated on the device
    X = ctx.rand_uniform((B, i), -1.0, 1.0, seed=100 + rank)
    dA = ctx.rand_normal((B, o), 0.0, 1.0, seed=200 + rank)
    W = ctx.rand_normal((o, i), 0.0, 0.5, seed=1)
    b = ctx.rand_normal((o,), 0.0, 0.5, seed=2)
    A = ctx.empty((B, o)); dX = ctx.empty((B, i))
    layout = dp.PackedLayout.for_layers([(o, i)])                            # [dW‖db]: one buffer, one all-reduce
    packed_t = torch.zeros(layout.numel, dtype=torch.float32, device=dev)
    packed = ctx.wrap_torch(packed_t)
    dWv, dbv = layout.views(packed)
    outs = (A, dX, dWv, dbv)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_nccl():
        nn.fflayer_fwd_grad(X, W, b, dA, out=outs)
        dp.allreduce_sum_(packed_t)

    # ---- data-parallel schedule.  Candidates (all verified against each other on this run's data before being timed):
    #   nccl      fwd, dW, dX, then ncclAllReduce([dW‖db]) on the same stream                                  (round-1 schedule)
    #   overlap-r tops_fflayer_step_dp: the all-reduce starts when dW/db are done and runs beside the dX GEMM, which leaves r SMs free
    #   fused     finished dW regions pushed from the dW GEMM epilogue with NVLS multimem.red (tops_fflayer_fwd_grad_mc)
    step, allreduce_kind, allreduce_trial = step_nccl, ("none (single GPU)" if world == 1 else "nccl"), None
    result_t = packed_t
    if world > 1:
        cands = {"nccl": (step_nccl, packed_t)}
        if args.allreduce in ("auto", "overlap"):
            for r in (0, 8, 16):
                ov = dp.OverlappedStep(ctx, layout, dev, reserve_sms=r)
                cands[f"overlap-{r}"] = ((lambda ov=ov: ov.step(X, W, b, dA, A=A, dX=dX)), ov.packed_t)
        if args.allreduce in ("auto", "fused"):
            try:
                fused = dp.FusedGradAllReduce(layout.numel, dev)
                local_grads = ctx.empty((layout.numel,))     # this rank's own [dW||db] (split-K accumulation target)

                def step_fused():
                    fused.begin()
                    nn.fflayer_fwd_grad_mc(X, W, b, dA, fused.multicast_ptr, out=(A, dX, local_grads))
                    fused.end()
                cands["fused"] = (step_fused, fused.local)
            except Exception as exc:   # no multicast support / symmetric memory unavailable
                if rank == 0:
                    print(f"bench.py: fused all-reduce unavailable ({exc}); skipped", file=sys.stderr)
        if args.allreduce != "auto":
            cands = {k: v for k, v in cands.items() if k.startswith(args.allreduce)} or {"nccl": (step_nccl, packed_t)}
        step_nccl(); torch.cuda.synchronize()
        want = packed_t.clone()
        okv = torch.ones(len(cands), device=dev)
        for k, (name, (fn, res)) in enumerate(cands.items()):
            try:
                fn(); fn(); torch.cuda.synchronize()
                err = float((res - want).norm() / want.norm())
                okv[k] = 1.0 if err < 1e-5 else 0.0
            except Exception as exc:
                okv[k] = 0.0
                if rank == 0:
                    print(f"bench.py: all-reduce candidate {name} failed: {exc}", file=sys.stderr)
        dist.all_reduce(okv, op=dist.ReduceOp.MIN)

        def quick(fn, n=8):
            for _ in range(2):
                fn()
            barrier()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            q0.record()
            for _ in range(n):
                fn()
            q1.record(); torch.cuda.synchronize()
            t = torch.tensor([q0.elapsed_time(q1) / n], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        allreduce_trial = {}
        for k, (name, (fn, res)) in enumerate(cands.items()):
            if float(okv[k].item()) == 1.0:
                allreduce_trial[name + "_ms_per_step"] = quick(fn)
        best = min(allreduce_trial, key=allreduce_trial.get)[: -len("_ms_per_step")]
        step, result_t = cands[best]
        allreduce_kind = {"nccl": "NCCL all-reduce of [dW||db] after the three GEMMs",
                          "fused": "finished dW regions pushed from the GEMM epilogue with NVLS multimem.red into symmetric memory"}.get(
                              best, f"tops_fflayer_step_dp: NCCL all-reduce of [dW||db] on a communication stream, started when dW/db are complete, overlapped with the dX GEMM ({best.split('-')[1]} SMs left free)")
        allreduce_kind += " — fastest of the candidates verified and timed on this run (allreduce_trial)"

    if world > 1:
        # the schedule trials above are ~60 back-to-back steps of preparation, not part of the benchmark: let the boards' power
        # management settle before the warm-up (multi-GPU boxes report sw_power_cap right after such a burst)
        torch.cuda.synchronize()
        time.sleep(1.0)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    n0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - n0
    # per-kernel durations for the roofline record: the same K steps once more, now with a CUDA event pair around every tagged
    # kernel (tops_profile_*).  Kept out of the timed region: the events sit between the launches and cost 1.2 % of a step
    # (tools/step_noprofile.py), i.e. a number taken with them would be a number taken under a profiler.
    if sampler:
        sampler.active = False
    torch.cuda.synchronize()
    time.sleep(1.0)          # same starting conditions as the timed region: the boards power-cap after ~50 back-to-back steps
    for _ in range(2):
        step()
    ctx.profile(True)
    for _ in range(args.steps):
        step()
    prof = ctx.profile_summary()
    ctx.profile(False)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = B * world / (ms_per_step * 1e-3)

    # ---- parity of THIS run: the timed inputs, the timed mode, against the fp64 oracle (rank 0; single-GPU semantics)
    parity = None
    if rank == 0:
        if world > 1:                      # A / dX of the timed step are per-rank; dW/db were summed over ranks: re-run locally
            nn.fflayer_fwd_grad(X, W, b, dA, out=(A, dX, None, None))
        parity = parity_check(ctx, nn, np, X, W, b, dA, A, dX)

    # ---- side measurement: the same step in single-pass TF32 (throughput mode; NOT the headline — its parity error is ~7e-4)
    side = None
    if args.precision in ("f16x3", "tf32bf16", "tf32x3") and not args.no_side:
        ctx.set_precision(tb.PREC_TF32)
        for _ in range(3):
            step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.profile(True)
        if sampler:
            sampler.active = True
        s0.record()
        for _ in range(args.steps):
            step()
        s1.record()
        barrier()
        if sampler:
            sampler.active = False
        sms = s0.elapsed_time(s1) / args.steps
        sprof = ctx.profile_summary()
        ctx.profile(False)
        ctx.set_precision(prec)
        if world > 1:
            t = torch.tensor([sms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sms = float(t.item())
        side = {"precision": "single-pass TF32 on tcgen05 (operands truncated to TF32: ~7e-4 relative error, fails the 1e-5 parity bar)",
                "ms_per_step": sms, "value": B * world / (sms * 1e-3), "unit": "samples/s",
                "per_kernel_ms": {k: v["ms"] / v["launches"] for k, v in sprof.items()}}

    # ---- e2e: host buffers in, gradient out, copies inside the timed region (same step, same sizes)
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((B, i), dtype=torch.float32, pin_memory=True); dAh = torch.empty((B, o), dtype=torch.float32, pin_memory=True)
        Xh.copy_(torch.from_numpy(X.numpy())); dAh.copy_(torch.from_numpy(dA.numpy()))
        gh = torch.empty(o * i + o, dtype=torch.float32, pin_memory=True)
        Xn, dAn, gn = Xh.numpy(), dAh.numpy(), gh.numpy()
        e2e_steps = max(3, min(args.steps, 10))

        def e2e_step():
            g = nn.fflayer_fwd_grad_host(ctx, Xn, W, b, dAn, grads_out=gn, allreduce=(lambda: dp.allreduce_sum_(packed_t)) if world > 1 else None,
                                         workspace=(A, dX, packed))
            return g
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler:
            sampler.active = True
        ee0.record()
        for _ in range(e2e_steps):
            e2e_step()
        ee1.record()
        barrier()
        if sampler:
            sampler.active = False
        wall = (time.perf_counter() - t0) * 1e3
        ems = max(ee0.elapsed_time(ee1), 0.0)
        ems = max(ems, wall) if ems == 0 else ems
        if world > 1:
            t = torch.tensor([ems], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        e2e = {"value": B * world / (ems / e2e_steps * 1e-3), "unit": "samples/s", "h2d_bytes_per_step": int(Xn.nbytes + dAn.nbytes),
               "d2h_bytes_per_step": int(gn.nbytes), "ms_per_step": ems / e2e_steps, "steps": e2e_steps,
               "what": "pinned host X,dA -> HBM, fwd+grad, [dW||db] -> pinned host, every step", "cpu_affinity": numa}

    # the sampler ran through the timed region, the TF32 side measurement and the e2e region: all of it is load
    clocks = sampler.stop() if sampler else None

    peaks, peaks_src = load_peaks()
    extra = None
    if not args.no_extra:
        extra = {"config4": config4_extra(ctx, tb, nn, dp, torch, dist, dev, rank, world, peaks)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tf32_peak_measured = None if args.no_extra else measure_tf32_peak(torch, dev)

    gemms = {k: v for k, v in prof.items() if k.startswith("gemm_")}
    dom = max(gemms, key=lambda k: gemms[k]["ms"] / max(1, gemms[k]["launches"]))
    dom_ms = gemms[dom]["ms"] / gemms[dom]["launches"]
    dom_flops = gemms[dom]["flops"] / gemms[dom]["launches"]
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    # fp32 configs are quoted against the TF32 tensor-core roofline.  MEASURED_PEAKS.json has no TF32 figure, so `peak` = measured
    # bf16 burst / 2 (TF32 dense is nominally half the bf16 rate on B200: 1.1 vs 2.25 PFLOP/s); the cuBLAS TF32 GEMM measured by this
    # very run (same recipe as the driver's bf16 number) is reported beside it with its own fraction.
    peak = peaks["bf16_tflops"] / 2.0
    passes = PASSES[args.precision]
    traffic, traffic_file = ncu_traffic_bytes(dom, args.precision)
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": (f"static: {traffic_file} (ncu --set full of this bench command, tools/run_ncu_gemm.sh; ncu cannot run inside a timed bench)" if traffic_file else None),
                "kernel": dom, "kernel_ms": dom_ms,
                "kernel_timing": "CUDA-event pairs around every tagged kernel (tops_profile_*) over a repetition of the K timed steps, right after them; kept out of the timed region because the events cost ~1.2 % of a step",
                "peak_tf32_measured": tf32_peak_measured, "frac_of_tf32_measured": (achieved / tf32_peak_measured) if tf32_peak_measured else None,
                "mma_flop_per_algorithmic_flop": passes, "tensor_pipe_frac": passes * achieved / peak,
                "note": ("parity modes spend more than one tensor-core pass per algorithmic FLOP, in TF32-pass equivalents: f16x3 = three fp16 passes at twice the "
                         "TF32 rate = 1.5 (frac capped at 2/3), tf32bf16 = 2 (cap 1/2), tf32x3 = 3 (cap 1/3); tensor_pipe_frac = passes*frac is the share of the "
                         "tensor-core peak the MMA stream itself reaches") if passes > 1 else "single pass",
                "peak_source": f"{peaks_src} bf16 burst {peaks['bf16_tflops']} TFLOP/s / 2 (TF32 = half the bf16 rate)",
                "per_kernel_ms": {k: v["ms"] / v["launches"] for k, v in prof.items()},
                "per_kernel_frac_of_peak": {k: (v["flops"] / v["launches"]) / (v["ms"] / v["launches"] * 1e-3) / 1e12 / peak for k, v in gemms.items()},
                "step_share": {k: v["ms"] / ms for k, v in prof.items()}}
    line = {"metric": "ffLayer fwd+grad samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "i": i, "o": o, "batch_per_gpu": B, "global_batch": B * world,
                       "precision": PREC_TEXT[args.precision],
                       "parallelism": f"dp{world} (batch-sharded, one all-reduce of [dW||db] per step)" if world > 1 else "single GPU",
                       "allreduce": allreduce_kind, "allreduce_trial": allreduce_trial,
                       "kernels": "5 per step: fp16-pair row split of X (+ max|dA|, max|W|), its fix-up pass (+ the pair of W and the layer scalars), then 3 tcgen05 cta_group::2 GEMMs (CTA pairs, 256x256 tiles): forward with fused bias/logistic/dZ-pair/db epilogue, split-K dW, dX" if args.precision == "f16x3" else "3 per step: tcgen05 cta_group::2 GEMMs (CTA pairs, 256x256 tiles) with fused bias/logistic/dZ/db epilogue, split-K dW, dX",
                       "l2": "inputs larger than L2: X and dA are 256 MiB each per step vs 126 MB L2"},
            "roofline": roofline, "parity": parity, "gpu_launches": int(launches), "clocks": clocks,
            "algorithmic_flop_per_step": 6.0 * B * i * o, "tflops_step": 6.0 * B * i * o / (ms_per_step * 1e-3) / 1e12}
    if side:
        gs = {k: v for k, v in side["per_kernel_ms"].items() if k.startswith("gemm_")}
        kdom = max(gs, key=gs.get)
        side["roofline"] = {"kernel": kdom, "kernel_ms": gs[kdom], "achieved": 2.0 * B * i * o / (gs[kdom] * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                            "frac": 2.0 * B * i * o / (gs[kdom] * 1e-3) / 1e12 / peak}
        line["throughput_mode"] = side
    if e2e:
        line["e2e"] = e2e
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        v, n, secs, threads = cpu_reference_rate(12.0, i, o)
        line["cpu_baseline"] = {"value": v, "unit": "samples/s", "cores": threads, "kind": "port",
                                "sample": f"{n} samples of the batch in {secs:.1f} s: per-sample hmatrix op sequence restated in NumPy/OpenBLAS fp32 (oracle.cpu_fflayer_step_reference), samples split over {threads} threads",
                                "variants": cpu_baseline_variants(i, o)}
    real_stdout.write(json.dumps(line) + "\n")
    real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["ok"]:
        print(f"bench.py: PARITY FAILURE on the timed inputs: {parity}", file=sys.stderr)
        sys.exit(3)


if __name__ == "__main__":
    main()
