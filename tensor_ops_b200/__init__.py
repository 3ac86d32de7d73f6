"""tensor_ops_b200 — a B200 (sm_100a) backend for the hot path of mstksg/tensor-ops: runTOp / gradTOp over ffLayer
networks, behind the reference's own Tensor / BLAS class surface.

Layout
  csrc/            hand-written CUDA (tcgen05 GEMM engine, bandwidth-bound kernels) and the C ABI (include/tops_b200.h)
  _lib.py          ctypes binding of libtops_b200.so (fails loudly when the library or the device is missing)
  tensor.py        CuTensor: `instance Tensor` with HBM storage         (src/TensorOps/Types.hs:52-109)
  expr.py          symbolic ElemT: reifies the host closures of liftT   (Types.hs:56-59, BLAS.hs:92-96)
  top.py           TOp, Category composition, routing, primitive TOps   (Types.hs:122-264, TOp.hs)
  nn.py            activations, losses, Network/ffLayer/genNet, fused batched entry points (Learn/NeuralNet*.hs)
  batched.py       BatchT: vmap-style `instance Tensor` with lazy outer products
  dp.py            data-parallel step: batch sharded over ranks, one all-reduce of [dW‖db]
"""
from . import _lib
from ._lib import (ACT_ID, ACT_LOGISTIC, ACT_SOFTMAX, BF16, F32, LOSS_CROSS_ENTROPY, LOSS_SQUARED_ERROR,
                   PREC_F16X3, PREC_FP32_SIMT, PREC_TF32, PREC_TF32X3, PREC_TF32_BF16X2, TopsError)
from .tensor import Context, CuTensor, default_context
from . import expr, top, nn, batched

__all__ = ["Context", "CuTensor", "default_context", "expr", "top", "nn", "batched", "TopsError",
           "F32", "BF16", "PREC_TF32X3", "PREC_TF32", "PREC_FP32_SIMT", "PREC_TF32_BF16X2", "PREC_F16X3", "ACT_ID", "ACT_LOGISTIC", "ACT_SOFTMAX",
           "LOSS_SQUARED_ERROR", "LOSS_CROSS_ENTROPY"]
