"""ctypes binding of libtops_b200.so (the C ABI declared in include/tops_b200.h).

The library is the product; this module only loads it and declares prototypes.  There is no fallback:
if the shared object is missing the import fails, and if no sm_100 device is present `Context()` raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtops_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found — build it with `make` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
        "tensor_ops_b200 has no CPU fallback.")

lib = C.CDLL(LIB_PATH)

c_ctx = C.c_void_p
c_buf = C.c_void_p
c_i64p = C.POINTER(C.c_int64)
c_bufp = C.POINTER(c_buf)

# name -> (restype, argtypes); every symbol include/tops_b200.h declares
PROTOTYPES = {
    "tops_init": (C.c_int, [C.c_int, C.POINTER(c_ctx)]),
    "tops_shutdown": (C.c_int, [c_ctx]),
    "tops_last_error": (C.c_char_p, [c_ctx]),
    "tops_sync": (C.c_int, [c_ctx]),
    "tops_set_stream": (C.c_int, [c_ctx, C.c_void_p]),
    "tops_set_precision": (C.c_int, [c_ctx, C.c_int]),
    "tops_get_precision": (C.c_int, [c_ctx]),
    "tops_launch_count": (C.c_int64, [c_ctx]),
    "tops_device_sm_count": (C.c_int, [c_ctx]),
    "tops_profile_enable": (C.c_int, [c_ctx, C.c_int]),
    "tops_profile_summary": (C.c_int, [c_ctx, C.c_char_p, C.c_size_t]),
    "tops_buf_alloc": (C.c_int, [c_ctx, C.c_int, C.c_int, c_i64p, c_bufp]),
    "tops_buf_wrap": (C.c_int, [c_ctx, C.c_void_p, C.c_int, C.c_int, c_i64p, c_bufp]),
    "tops_buf_view": (C.c_int, [c_ctx, c_buf, C.c_int64, C.c_int, c_i64p, c_bufp]),
    "tops_buf_retain": (C.c_int, [c_buf]),
    "tops_buf_release": (C.c_int, [c_buf]),
    "tops_buf_rank": (C.c_int, [c_buf]),
    "tops_buf_dims": (C.c_int, [c_buf, c_i64p]),
    "tops_buf_dtype": (C.c_int, [c_buf]),
    "tops_buf_numel": (C.c_int64, [c_buf]),
    "tops_buf_data": (C.c_void_p, [c_buf]),
    "tops_host_alloc": (C.c_int, [C.c_size_t, C.c_int, C.POINTER(C.c_void_p)]),
    "tops_host_free": (C.c_int, [C.c_void_p]),
    "tops_upload": (C.c_int, [c_ctx, c_buf, C.c_void_p, C.c_size_t]),
    "tops_download": (C.c_int, [c_ctx, c_buf, C.c_void_p, C.c_size_t]),
    "tops_fill": (C.c_int, [c_ctx, c_buf, C.c_double]),
    "tops_rand_normal": (C.c_int, [c_ctx, c_buf, C.c_double, C.c_double, C.c_uint64]),
    "tops_rand_uniform": (C.c_int, [c_ctx, c_buf, C.c_double, C.c_double, C.c_uint64]),
    "tops_cast": (C.c_int, [c_ctx, c_buf, C.c_int, c_bufp]),
    "tops_axpy": (C.c_int, [c_ctx, C.c_double, c_buf, c_buf, c_bufp]),
    "tops_dot": (C.c_int, [c_ctx, c_buf, c_buf, c_bufp]),
    "tops_ger": (C.c_int, [c_ctx, c_buf, c_buf, c_bufp]),
    "tops_gemv": (C.c_int, [c_ctx, C.c_double, c_buf, c_buf, C.c_double, c_buf, c_bufp]),
    "tops_gemm": (C.c_int, [c_ctx, C.c_double, c_buf, c_buf, C.c_double, c_buf, c_bufp]),
    "tops_scale": (C.c_int, [c_ctx, C.c_double, c_buf, c_bufp]),
    "tops_add": (C.c_int, [c_ctx, c_buf, c_buf, c_bufp]),
    "tops_index": (C.c_int, [c_ctx, c_buf, c_i64p, C.POINTER(C.c_double)]),
    "tops_index_row": (C.c_int, [c_ctx, c_buf, C.c_int64, c_bufp]),
    "tops_transp": (C.c_int, [c_ctx, c_buf, c_bufp]),
    "tops_eye": (C.c_int, [c_ctx, C.c_int64, c_bufp]),
    "tops_trace": (C.c_int, [c_ctx, c_buf, c_bufp]),
    "tops_diag": (C.c_int, [c_ctx, C.c_int, c_buf, c_bufp]),
    "tops_get_diag": (C.c_int, [c_ctx, c_buf, c_bufp]),
    "tops_sum": (C.c_int, [c_ctx, c_buf, c_bufp]),
    "tops_lift": (C.c_int, [c_ctx, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int, c_bufp, C.c_int, c_i64p, c_bufp]),
    "tops_lift_catalogue_hits": (C.c_int64, []),
    "tops_gmul": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, c_buf, c_buf, c_bufp]),
    "tops_gmul_sum_rows": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, c_buf, c_buf, c_bufp]),
    "tops_gmul_sum_rows_vjp": (C.c_int, [c_ctx, C.c_int, C.c_int, C.c_int, c_buf, c_buf, c_buf, c_bufp, c_bufp]),
    "tops_sum_t": (C.c_int, [c_ctx, C.c_int, c_bufp, c_bufp]),
    "tops_sum_rows": (C.c_int, [c_ctx, c_buf, c_bufp]),
    "tops_broadcast_rows": (C.c_int, [c_ctx, C.c_int64, c_buf, c_bufp]),
    "tops_map_rows_softmax": (C.c_int, [c_ctx, c_buf, c_bufp]),
    "tops_fflayer_fwd": (C.c_int, [c_ctx, c_buf, c_buf, c_buf, C.c_int, c_bufp]),
    "tops_fflayer_grad": (C.c_int, [c_ctx, c_buf, c_buf, c_buf, C.c_int, c_buf, c_buf, c_bufp, c_bufp, c_bufp]),
    "tops_fflayer_fwd_grad": (C.c_int, [c_ctx, c_buf, c_buf, c_buf, C.c_int, c_buf, c_bufp, c_bufp, c_bufp, c_bufp]),
    "tops_mlp_fwd_grad": (C.c_int, [c_ctx, C.c_int, c_bufp, c_bufp, C.POINTER(C.c_int), C.c_int, c_buf, c_buf, c_bufp, c_bufp, c_bufp, c_bufp, c_bufp]),
    "tops_mlp_fwd": (C.c_int, [c_ctx, C.c_int, c_bufp, c_bufp, C.POINTER(C.c_int), c_buf, c_bufp]),
    "tops_fflayer_fwd_grad_host": (C.c_int, [c_ctx, C.c_void_p, C.c_void_p, C.c_int64, c_buf, c_buf, C.c_int, C.c_int, c_bufp, c_bufp, c_bufp, C.c_void_p]),
    "tops_fflayer_fwd_grad_mc": (C.c_int, [c_ctx, c_buf, c_buf, c_buf, C.c_int, c_buf, c_bufp, c_bufp, c_bufp, C.c_void_p]),
    "tops_fflayer_step_dp": (C.c_int, [c_ctx, c_buf, c_buf, c_buf, C.c_int, c_buf, c_bufp, c_bufp, c_bufp, C.c_void_p, C.c_int]),
    "tops_graph_begin": (C.c_int, [c_ctx, C.c_size_t, C.POINTER(C.c_void_p)]),
    "tops_graph_end": (C.c_int, [c_ctx, C.c_void_p]),
    "tops_graph_launch": (C.c_int, [c_ctx, C.c_void_p]),
    "tops_graph_kernel_count": (C.c_int64, [C.c_void_p]),
    "tops_graph_destroy": (C.c_int, [c_ctx, C.c_void_p]),
    "tops_copy": (C.c_int, [c_ctx, c_buf, c_buf]),
    "tops_event_create": (C.c_int, [c_ctx, C.POINTER(C.c_void_p)]),
    "tops_event_destroy": (C.c_int, [c_ctx, C.c_void_p]),
    "tops_stream_wait_event": (C.c_int, [c_ctx, C.c_void_p, C.c_void_p]),
    "tops_sgd_step": (C.c_int, [c_ctx, C.c_int, c_bufp, c_bufp, C.c_double, c_bufp]),
}

for _name, (_res, _args) in PROTOTYPES.items():
    _fn = getattr(lib, _name)          # AttributeError here == the .so does not export a declared symbol
    _fn.restype = _res
    _fn.argtypes = _args

# enums (include/tops_b200.h)
F32, BF16 = 0, 1
PREC_TF32X3, PREC_TF32, PREC_FP32_SIMT, PREC_TF32_BF16X2, PREC_F16X3 = 0, 1, 2, 3, 4
ACT_ID, ACT_LOGISTIC, ACT_SOFTMAX = 0, 1, 2
LOSS_NONE, LOSS_SQUARED_ERROR, LOSS_CROSS_ENTROPY = 0, 1, 2
OK = 0
STATUS_NAMES = {1: "INVALID", 2: "SHAPE", 3: "CUDA", 4: "OOM", 5: "UNSUPPORTED", 6: "NO_DEVICE"}

OPCODES = dict(VAR=0, CONST=1, ADD=2, SUB=3, MUL=4, DIV=5, NEG=6, EXP=7, LOG=8, RECIP=9, SQRT=10, TANH=11,
               ABS=12, SIGNUM=13, MAX=14, MIN=15, POW=16, LOGISTIC=17, SIN=18, COS=19)


class TopsError(RuntimeError):
    """Non-zero status from the C ABI (the Haskell shim turns the same codes into `error`)."""

    def __init__(self, code: int, msg: str):
        super().__init__(f"tops error {code} ({STATUS_NAMES.get(code, '?')}): {msg}")
        self.code = code
