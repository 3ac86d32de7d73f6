"""`CuTensor`: the device-HBM instance of tensor-ops' `class Tensor` (src/TensorOps/Types.hs:52-109).

Storage lives in B200 HBM behind an opaque ref-counted `tops_buf`; every class method is one call through the
C ABI (include/tops_b200.h) that enqueues sm_100a kernels on the context's stream.  Data crosses PCIe only at the
reference's own observation points: `fromList`/`generateA` (upload) and `toList`/`(!)`/`unScalar` (download).
Method names and argument order follow the reference so that TOp closures read the same.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import _lib as L
from . import expr as E

_NP_OF = {L.F32: np.float32}


class Context:
    """One CUDA context/stream of the library (`tops_init`).  Raises if there is no sm_100 device: no fallback."""

    def __init__(self, device: int = 0):
        h = L.c_ctx()
        rc = L.lib.tops_init(device, C.byref(h))
        if rc != L.OK:
            raise L.TopsError(rc, "tops_init failed: no usable sm_100 (B200) device — tensor_ops_b200 has no CPU fallback")
        self.h = h
        self.device = device
        self._fin = weakref.finalize(self, L.lib.tops_shutdown, h)

    def check(self, rc: int):
        if rc != L.OK:
            raise L.TopsError(rc, L.lib.tops_last_error(self.h).decode())

    def sync(self): self.check(L.lib.tops_sync(self.h))
    def set_stream(self, cuda_stream: Optional[int]): self.check(L.lib.tops_set_stream(self.h, C.c_void_p(cuda_stream or 0)))
    def set_precision(self, p: int): self.check(L.lib.tops_set_precision(self.h, p))
    def precision(self) -> int: return L.lib.tops_get_precision(self.h)
    def launch_count(self) -> int: return L.lib.tops_launch_count(self.h)
    def sm_count(self) -> int: return L.lib.tops_device_sm_count(self.h)
    def profile(self, on: bool): self.check(L.lib.tops_profile_enable(self.h, int(on)))

    def record(self, arena_bytes: int = 64 << 20) -> "Graph":
        """`with ctx.record() as g: ...` records every library call of the block into a graph (tops_graph_*); `g.launch()` replays it."""
        return Graph(self, arena_bytes)

    def profile_summary(self) -> dict:
        """Per-tag device times of the launches since profiling was enabled (synchronises)."""
        import json
        buf = C.create_string_buffer(8192)
        self.check(L.lib.tops_profile_summary(self.h, buf, len(buf)))
        return json.loads(buf.value.decode())

    # ---- construction
    def empty(self, dims: Sequence[int], dtype: int = L.F32) -> "CuTensor":
        d = (C.c_int64 * max(1, len(dims)))(*dims)
        b = L.c_buf()
        self.check(L.lib.tops_buf_alloc(self.h, dtype, len(dims), d, C.byref(b)))
        return CuTensor(self, b)

    def full(self, dims, v: float, dtype: int = L.F32) -> "CuTensor":
        t = self.empty(dims, dtype)
        self.check(L.lib.tops_fill(self.h, t.b, float(v)))
        return t

    def from_numpy(self, a: np.ndarray) -> "CuTensor":
        """`fromList` / `generateA` (Tensor.hs:187-205): host -> HBM."""
        a = np.ascontiguousarray(a, dtype=np.float32)
        t = self.empty(a.shape, L.F32)
        self.check(L.lib.tops_upload(self.h, t.b, a.ctypes.data_as(C.c_void_p), a.nbytes))
        self.sync()   # `a` may be a temporary
        return t

    def rand_normal(self, dims, mean=0.0, sd=1.0, seed=0) -> "CuTensor":
        t = self.empty(dims)
        self.check(L.lib.tops_rand_normal(self.h, t.b, mean, sd, seed))
        return t

    def rand_uniform(self, dims, lo=0.0, hi=1.0, seed=0) -> "CuTensor":
        t = self.empty(dims)
        self.check(L.lib.tops_rand_uniform(self.h, t.b, lo, hi, seed))
        return t

    def host_empty(self, shape, write_combined: bool = False) -> np.ndarray:
        """A float32 NumPy array over page-locked host memory (tops_host_alloc): the staging buffer for `upload` / the host-buffer
        entry points.  write_combined=True: faster for the GPU to read over PCIe, very slow for the CPU to read back."""
        n = int(np.prod(shape)) if len(shape) else 1
        p = C.c_void_p()
        self.check(L.lib.tops_host_alloc(C.c_size_t(4 * max(n, 1)), int(write_combined), C.byref(p)))
        buf = (C.c_float * max(n, 1)).from_address(p.value)
        # the ARRAY owns the block: arr.base -> memoryview -> buf -> _block, so the pinned memory lives exactly as long as any
        # array (or view of it) does, independently of this Context object
        buf._block = _HostBlock(p.value)
        return np.frombuffer(buf, dtype=np.float32, count=n).reshape(shape)

    def wrap(self, device_ptr: int, dims, dtype: int = L.F32, keepalive=None) -> "CuTensor":
        """Non-owning view of caller-allocated device memory (e.g. a torch tensor's storage)."""
        d = (C.c_int64 * max(1, len(dims)))(*dims)
        b = L.c_buf()
        self.check(L.lib.tops_buf_wrap(self.h, C.c_void_p(device_ptr), dtype, len(dims), d, C.byref(b)))
        t = CuTensor(self, b)
        t._keep = keepalive
        return t

    def wrap_torch(self, t) -> "CuTensor":
        import torch
        assert t.is_cuda and t.is_contiguous()
        dt = {torch.float32: L.F32, torch.bfloat16: L.BF16}[t.dtype]
        return self.wrap(t.data_ptr(), tuple(t.shape), dt, keepalive=t)


class Graph:
    """A recorded sequence of library calls replayed as one CUDA graph launch (tops_graph_begin/end/launch): the deferred,
    on-device evaluation of a composed TOp pipeline — forward, reverse sweep, parameter update — without per-method launch latency.
    Tensors created inside the `with` block live in the graph's arena: keep the Graph alive while they are used."""

    def __init__(self, ctx: Context, arena_bytes: int = 64 << 20):
        self.ctx, self.arena_bytes, self.h = ctx, int(arena_bytes), None

    def __enter__(self):
        h = C.c_void_p()
        self.ctx.check(L.lib.tops_graph_begin(self.ctx.h, self.arena_bytes, C.byref(h)))
        self.h = h
        return self

    def __exit__(self, et, ev, tb):
        if et is not None:               # abandon the recording
            L.lib.tops_graph_destroy(self.ctx.h, self.h)
            self.h = None
            return False
        self.ctx.check(L.lib.tops_graph_end(self.ctx.h, self.h))
        return False

    def launch(self):
        self.ctx.check(L.lib.tops_graph_launch(self.ctx.h, self.h))

    def kernel_count(self) -> int:
        return int(L.lib.tops_graph_kernel_count(self.h))

    def close(self):
        if self.h is not None:
            self.ctx.check(L.lib.tops_graph_destroy(self.ctx.h, self.h))
            self.h = None


_default: Optional[Context] = None


class _HostBlock:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            L.lib.tops_host_free(C.c_void_p(self.ptr))
        except Exception:
            pass


def default_context() -> Context:
    global _default
    if _default is None:
        _default = Context(0)
    return _default


_prog_cache = {}


def _compiled(e: E.Expr):
    k = e.key()
    hit = _prog_cache.get(k)
    if hit is None:
        code, consts = E.compile_expr(e)
        hit = ((C.c_int32 * len(code))(*code), len(code), (C.c_float * max(1, len(consts)))(*consts), len(consts))
        _prog_cache[k] = hit
    return hit


def _bufs(xs: Sequence["CuTensor"]):
    return (L.c_buf * max(1, len(xs)))(*[x.b for x in xs])


class CuTensor:
    """instance Tensor CuTensor — `ElemT CuTensor` is the symbolic `expr.Expr` when lifting, fp32 on the device."""

    __slots__ = ("ctx", "b", "_fin", "_keep", "__weakref__")

    def __init__(self, ctx: Context, b):
        self.ctx, self.b, self._keep = ctx, b, None
        self._fin = weakref.finalize(self, L.lib.tops_buf_release, b)   # ForeignPtr finalizer in the Haskell shim

    # ---- shape / host access (the reference's observation points)
    @property
    def shape(self):
        r = L.lib.tops_buf_rank(self.b)
        d = (C.c_int64 * max(1, r))()
        L.lib.tops_buf_dims(self.b, d)
        return tuple(d[i] for i in range(r))

    @property
    def dtype(self) -> int: return L.lib.tops_buf_dtype(self.b)
    @property
    def data_ptr(self) -> int: return L.lib.tops_buf_data(self.b)

    def numpy(self) -> np.ndarray:
        """`toList` (Tensor.hs:262-266): HBM -> host, synchronises."""
        if self.dtype == L.BF16:
            raw = np.empty(self.shape, dtype=np.uint16)
            self.ctx.check(L.lib.tops_download(self.ctx.h, self.b, raw.ctypes.data_as(C.c_void_p), raw.nbytes))
            return (raw.astype(np.uint32) << 16).view(np.float32)
        out = np.empty(self.shape, dtype=np.float32)
        self.ctx.check(L.lib.tops_download(self.ctx.h, self.b, out.ctypes.data_as(C.c_void_p), out.nbytes))
        return out

    def upload(self, host: np.ndarray, sync: bool = False) -> "CuTensor":
        """host -> this tensor's HBM (async on the context's stream when `host` is pinned)."""
        assert host.flags["C_CONTIGUOUS"]
        self.ctx.check(L.lib.tops_upload(self.ctx.h, self.b, host.ctypes.data_as(C.c_void_p), host.nbytes))
        if sync:
            self.ctx.sync()
        return self

    def copy_from(self, src: "CuTensor") -> "CuTensor":
        """this tensor's storage <- src (device to device, tops_copy): how a recorded step publishes new state in place."""
        self.ctx.check(L.lib.tops_copy(self.ctx.h, self.b, src.b))
        return self

    def download_into(self, host: np.ndarray) -> np.ndarray:
        """HBM -> an existing (ideally pinned) host array; synchronises."""
        assert host.flags["C_CONTIGUOUS"]
        self.ctx.check(L.lib.tops_download(self.ctx.h, self.b, host.ctypes.data_as(C.c_void_p), host.nbytes))
        return host

    def index(self, idx: Sequence[int]) -> float:
        """`(!)` (Types.hs:107-109)."""
        v = C.c_double()
        i = (C.c_int64 * max(1, len(idx)))(*idx)
        self.ctx.check(L.lib.tops_index(self.ctx.h, self.b, i, C.byref(v)))
        return v.value

    def unScalar(self) -> float:
        """`TT.unScalar` (Tensor.hs:268-273)."""
        return self.index(())

    def _new(self, fn, *args) -> "CuTensor":
        out = L.c_buf()
        self.ctx.check(fn(self.ctx.h, *args, C.byref(out)))
        return CuTensor(self.ctx, out)

    # ---- class Tensor (Types.hs:52-109), as static methods so TOp closures receive the "dictionary" `T`
    @staticmethod
    def liftT(f: Callable, xs: Sequence["CuTensor"], like: Optional["CuTensor"] = None) -> "CuTensor":
        """`liftT` (Types.hs:56-59): `f` is applied once to symbolic variables and shipped as bytecode."""
        n = len(xs)
        e = E.trace(f, n)
        ref = xs[0] if n else like
        prog, plen, consts, nc = _compiled(e)
        shape = ref.shape
        d = (C.c_int64 * max(1, len(shape)))(*shape)
        out = L.c_buf()
        ref.ctx.check(L.lib.tops_lift(ref.ctx.h, prog, plen, consts, nc, n, _bufs(xs), len(shape), d, C.byref(out)))
        return CuTensor(ref.ctx, out)

    @staticmethod
    def gmul(lM: int, lO: int, lN: int, x: "CuTensor", y: "CuTensor") -> "CuTensor":
        """`gmul` (Types.hs:60-66): one tensor-core GEMM on flat storage instead of BTensor's rank dispatch."""
        return x._new(L.lib.tops_gmul, lM, lO, lN, x.b, y.b)

    @staticmethod
    def gmulSumRows(lM: int, lO: int, lN: int, x: "CuTensor", y: "CuTensor") -> "CuTensor":
        """`gmul lM lO lN >>> sumRows` fused (tops_gmul_sum_rows): one pass over x, no [A, ...] intermediate."""
        return x._new(L.lib.tops_gmul_sum_rows, lM, lO, lN, x.b, y.b)

    @staticmethod
    def gmulSumRowsVJP(lM: int, lO: int, lN: int, x: "CuTensor", y: "CuTensor", ct: "CuTensor"):
        dx, dy = L.c_buf(), L.c_buf()
        x.ctx.check(L.lib.tops_gmul_sum_rows_vjp(x.ctx.h, lM, lO, lN, x.b, y.b, ct.b, C.byref(dx), C.byref(dy)))
        return CuTensor(x.ctx, dx), CuTensor(x.ctx, dy)

    @staticmethod
    def sumT(xs: Sequence["CuTensor"]) -> "CuTensor":
        """`sumT` (Types.hs:69)."""
        if len(xs) == 1:
            return xs[0]
        return xs[0]._new(L.lib.tops_sum_t, len(xs), _bufs(xs))

    @staticmethod
    def scaleT(a: float, x: "CuTensor") -> "CuTensor":
        """`scaleT` (Types.hs:70)."""
        return x._new(L.lib.tops_scale, float(a), x.b)

    @staticmethod
    def transp(x: "CuTensor") -> "CuTensor":
        """`transp` (Types.hs:71-73): full axis reversal; O(1) view for rank <= 2."""
        return x._new(L.lib.tops_transp, x.b)

    @staticmethod
    def mapRows(lN: int, f: Callable[["CuTensor"], "CuTensor"], x: "CuTensor") -> "CuTensor":
        """`mapRows` (Types.hs:77-81) with a host-level function on sub-tensors (views; no host round trip of data)."""
        shape = x.shape
        lead = shape[:lN]
        rows = int(np.prod(lead)) if lead else 1
        sub = shape[lN:]
        flat = x.reshape((rows,) + tuple(sub))
        outs = [f(flat.row(r)) for r in range(rows)]
        out = x.ctx.empty(shape)
        n_sub = int(np.prod(sub)) if sub else 1
        for r, o in enumerate(outs):
            view = out.view(r * n_sub, sub)
            v = view.b
            x.ctx.check(L.lib.tops_axpy(x.ctx.h, 1.0, o.b, None, C.byref(v)))
        return out

    @staticmethod
    def sumRows(x: "CuTensor") -> "CuTensor":
        """`sumRows` (Types.hs:82-84)."""
        return x._new(L.lib.tops_sum_rows, x.b)

    @staticmethod
    def broadcastRows(n: int, row: "CuTensor") -> "CuTensor":
        """`mapRows (LS LZ) (const row)` — the VJP of sumRows (TOp.hs:151-159) without a per-row loop."""
        return row._new(L.lib.tops_broadcast_rows, n, row.b)

    @staticmethod
    def diag(rank: int, v: "CuTensor") -> "CuTensor":
        """`diag` (Types.hs:85-88)."""
        return v._new(L.lib.tops_diag, rank, v.b)

    @staticmethod
    def getDiag(x: "CuTensor") -> "CuTensor":
        """`getDiag` (Types.hs:89-92)."""
        return x._new(L.lib.tops_get_diag, x.b)

    @staticmethod
    def konst(shape: Sequence[int], v: float, like: "CuTensor") -> "CuTensor":
        """`TT.konst` (Tensor.hs:49-54)."""
        ctx = like.ctx if like is not None else default_context()
        return ctx.full(shape, v)

    # ---- BLAS class surface (BLAS.hs:90-173) for callers that want it directly
    def axpy(self, alpha: float, y: Optional["CuTensor"] = None): return self._new(L.lib.tops_axpy, float(alpha), self.b, y.b if y is not None else None)
    def dot(self, y: "CuTensor"): return self._new(L.lib.tops_dot, self.b, y.b)
    def ger(self, y: "CuTensor"): return self._new(L.lib.tops_ger, self.b, y.b)
    def gemv(self, x: "CuTensor", alpha=1.0, beta=0.0, y: Optional["CuTensor"] = None): return self._new(L.lib.tops_gemv, float(alpha), self.b, x.b, float(beta), y.b if y is not None else None)
    def gemm(self, b: "CuTensor", alpha=1.0, beta=0.0, c: Optional["CuTensor"] = None): return self._new(L.lib.tops_gemm, float(alpha), self.b, b.b, float(beta), c.b if c is not None else None)
    def trace(self): return self._new(L.lib.tops_trace, self.b)
    def sum(self): return self._new(L.lib.tops_sum, self.b)
    def row(self, i: int): return self._new(L.lib.tops_index_row, self.b, int(i))
    def cast(self, dtype: int): return self._new(L.lib.tops_cast, self.b, dtype)

    def view(self, offset: int, dims: Sequence[int]) -> "CuTensor":
        d = (C.c_int64 * max(1, len(dims)))(*dims)
        out = L.c_buf()
        self.ctx.check(L.lib.tops_buf_view(self.ctx.h, self.b, int(offset), len(dims), d, C.byref(out)))
        t = CuTensor(self.ctx, out)
        t._keep = self
        return t

    def reshape(self, dims: Sequence[int]) -> "CuTensor":
        return self.view(0, dims)

    def __add__(self, o: "CuTensor"): return self._new(L.lib.tops_add, self.b, o.b)
    def __repr__(self): return f"CuTensor(shape={self.shape}, dtype={'bf16' if self.dtype else 'f32'})"


def eye(ctx: Context, n: int) -> CuTensor:
    out = L.c_buf()
    ctx.check(L.lib.tops_eye(ctx.h, n, C.byref(out)))
    return CuTensor(ctx, out)
