/* tops_b200.h — C ABI of libtops_b200.so: a B200 (sm_100a) backend for the hot path of mstksg/tensor-ops,
 * i.e. evaluating a composed TOp (runTOp / gradTOp) over ffLayer networks.
 *
 * The reference has NO C ABI and NO FFI (grep for `foreign import` over /root/reference is empty); its plugin
 * boundary is two Haskell type-class dictionaries:
 *     class BLAS   (src/TensorOps/BLAS.hs:90-173)   — what Backend/BTensor.hs dispatches onto
 *     class Tensor (src/TensorOps/Types.hs:52-109)  — what every TOp closure is written against
 * Each entry point below names the class method (file:line) it is the device-side body of.  A Haskell
 * `instance BLAS CuMat` / `instance Tensor CuTensor` binds them with `foreign import ccall`.
 *
 * Conventions
 *   - Every function returns 0 (TOPS_OK) or a tops_status code; it never throws, aborts or exits.
 *     `tops_last_error(ctx)` returns the message for the last failure on that context.
 *   - Tensors are immutable, ref-counted device buffers (`tops_buf`), row-major, first index outermost, exactly
 *     the index order of the reference's type-level dimension lists.  Methods return NEW buffers through
 *     `tops_buf** out`; if `*out` is non-NULL on entry it must be a buffer of the right shape and is
 *     written in place (lets callers pre-allocate / pack outputs, e.g. [dW‖db] for one all-reduce).
 *   - All work is enqueued on the context's stream; nothing synchronises except tops_sync, tops_download
 *     and tops_index (the reference's observation points `(!)`, `toList`, `indexB`).
 *   - There is NO CPU fallback: if the CUDA device or the sm_100a kernels are unavailable, tops_init fails.
 *   - Thread safety: a context may be used from several host threads; calls are serialised by a mutex.
 */
#ifndef TOPS_B200_H
#define TOPS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tops_ctx tops_ctx;
typedef struct tops_buf tops_buf;

typedef enum {
    TOPS_OK = 0,
    TOPS_ERR_INVALID = 1,      /* bad argument / NULL */
    TOPS_ERR_SHAPE = 2,        /* shapes do not conform */
    TOPS_ERR_CUDA = 3,         /* CUDA runtime / launch failure (message has the CUDA error) */
    TOPS_ERR_OOM = 4,
    TOPS_ERR_UNSUPPORTED = 5,
    TOPS_ERR_NO_DEVICE = 6     /* no sm_100 device: the library refuses to run (no fallback) */
} tops_status;

typedef enum { TOPS_F32 = 0, TOPS_BF16 = 1 } tops_dtype;

/* How fp32 GEMM-class work (gemm, gmul with |os|>=1 on matrices, the fused ffLayer paths) uses the tensor cores. */
typedef enum {
    TOPS_PREC_TF32X3 = 0,  /* 3-pass hi/lo TF32 split on tcgen05 (hi*hi + lo*hi + hi*lo), fp32-grade accuracy (~1.5e-6 rel) */
    TOPS_PREC_TF32 = 1,    /* single TF32 pass on tcgen05 (~1e-3 rel), the throughput mode                   */
    TOPS_PREC_FP32_SIMT = 2, /* CUDA-core FFMA kernel (exact fp32 products); also what un-TMA-able strides use */
    TOPS_PREC_TF32_BF16X2 = 3, /* hi*hi in TF32 + the two first-order corrections bf16(lo)*bf16(x) as bf16 MMAs (half the cost of a
                                 TF32 pass each): fp32-grade accuracy (~1.4e-6 rel) at 2 tensor-core passes instead of 3 */
    TOPS_PREC_F16X3 = 4       /* default: every fp32 operand is stored once as an fp16 PAIR (hi, lo) of x * 2^k (22 bits of mantissa, k per
                                 tensor or per row) and the product is three fp16 tcgen05 passes hi*hi + lo*hi + hi*lo with fp32
                                 accumulation: ~6e-7 rel at 1.5 TF32-pass equivalents (fp16 MMAs run at twice the TF32 rate) */
} tops_precision;

/* Routing that does not depend on the mode: fp32 products with M, N or K <= 16 (an MLP's output layer) are HBM-bound and run on
 * streaming CUDA-core kernels in plain fp32 (exact products) in every mode but TOPS_PREC_FP32_SIMT, which keeps its one reference kernel.
 * Environment overrides read by tops_init (diagnostics, not API): TOPS_SKINNY=0 keeps those products on the tensor-core path;
 * TOPS_F16X3_CHUNK / TOPS_F16X3_FWD_CHUNK / TOPS_F16X3_FWD_HEAD = k-blocks per TMEM accumulation chunk (accuracy/speed trade-off of
 * TOPS_PREC_F16X3); TOPS_GEMM_DEBUG bit 0 disables the specialised epilogue code paths of the tcgen05 GEMM. */
typedef enum { TOPS_ACT_ID = 0, TOPS_ACT_LOGISTIC = 1, TOPS_ACT_SOFTMAX = 2 } tops_act;
typedef enum { TOPS_LOSS_NONE = 0, TOPS_LOSS_SQUARED_ERROR = 1, TOPS_LOSS_CROSS_ENTROPY = 2 } tops_loss;

#define TOPS_MAX_RANK 8

/* ------------------------------------------------------------------ lifecycle */
int tops_init(int device, tops_ctx** ctx);
int tops_shutdown(tops_ctx* ctx);
const char* tops_last_error(tops_ctx* ctx);
int tops_sync(tops_ctx* ctx);
/* Run on an externally owned CUDA stream (e.g. torch's current stream); NULL restores the context's own stream. */
int tops_set_stream(tops_ctx* ctx, void* cuda_stream);
int tops_set_precision(tops_ctx* ctx, int precision);
int tops_get_precision(tops_ctx* ctx);
/* number of kernels this library has launched on the context so far (bench.py's gpu_launches) */
int64_t tops_launch_count(tops_ctx* ctx);
int tops_device_sm_count(tops_ctx* ctx);
/* Per-kernel device timing for bench.py's roofline: when enabled, tagged launches (gemm_fwd / gemm_dX / gemm_dW /
 * col_sums_db / gemm / gmul) are bracketed by CUDA events on the context's stream.  tops_profile_summary synchronises,
 * writes {"tag": {"launches", "ms", "flops", "bytes"}} as JSON into `json_out` and clears the records. */
int tops_profile_enable(tops_ctx* ctx, int on);
int tops_profile_summary(tops_ctx* ctx, char* json_out, size_t cap);

/* ------------------------------------------------------------------ storage
 * replaces: hmatrix Storable Vector/Matrix allocation inside every HMat method (BLAS/HMat.hs:37-39),
 *           `generateA` / `genRand` / `fromList` / `toList` / `(!)` (Types.hs:93-109, Backend/BTensor.hs:838-858). */
int tops_buf_alloc(tops_ctx* ctx, int dtype, int rank, const int64_t* dims, tops_buf** out);
/* non-owning view of device memory the caller allocated (torch tensor, cudaMalloc); caller keeps it alive */
int tops_buf_wrap(tops_ctx* ctx, void* device_ptr, int dtype, int rank, const int64_t* dims, tops_buf** out);
/* contiguous sub-range view: `offset` elements into `parent`, new shape; keeps parent alive */
int tops_buf_view(tops_ctx* ctx, tops_buf* parent, int64_t offset, int rank, const int64_t* dims, tops_buf** out);
int tops_buf_retain(tops_buf* b);
int tops_buf_release(tops_buf* b);     /* stream-ordered free when the count reaches 0 */
int tops_buf_rank(const tops_buf* b);
int tops_buf_dims(const tops_buf* b, int64_t* dims_out);   /* writes rank entries */
int tops_buf_dtype(const tops_buf* b);
int64_t tops_buf_numel(const tops_buf* b);
void* tops_buf_data(const tops_buf* b);                     /* device pointer */
/* page-locked host staging memory for upload / download / the host-buffer entry points (full PCIe rate, asynchronous copies);
 * write_combined = 1 asks for write-combined pages: faster for the device to read, slow for the CPU to READ (fill it, don't read it) */
int tops_host_alloc(size_t bytes, int write_combined, void** out);
int tops_host_free(void* p);
int tops_upload(tops_ctx* ctx, tops_buf* dst, const void* host, size_t bytes);      /* H2D, async if host is pinned */
int tops_download(tops_ctx* ctx, const tops_buf* src, void* host, size_t bytes);    /* D2H + stream sync */
int tops_fill(tops_ctx* ctx, tops_buf* dst, double value);                          /* TT.konst, Tensor.hs:49-54 */
int tops_rand_normal(tops_ctx* ctx, tops_buf* dst, double mean, double stddev, uint64_t seed);   /* genRand (normalDistr ..), FeedForward.hs:206-207 */
int tops_rand_uniform(tops_ctx* ctx, tops_buf* dst, double lo, double hi, uint64_t seed);        /* genRand (uniformDistr ..), Dots.hs:63 */
int tops_cast(tops_ctx* ctx, const tops_buf* x, int dtype, tops_buf** out);

/* ------------------------------------------------------------------ class BLAS, method for method (BLAS.hs:90-173;
 * reference bodies BLAS/HMat.hs:103-231).  Vectors are rank-1, matrices rank-2 row-major.  y/c operands may be NULL
 * (the reference's `Maybe`). */
int tops_axpy(tops_ctx*, double alpha, const tops_buf* x, const tops_buf* y, tops_buf** out);                 /* BLAS.hs:97-101  HMat.hs:135-139 */
int tops_dot(tops_ctx*, const tops_buf* x, const tops_buf* y, tops_buf** out_scalar);                         /* BLAS.hs:102-104 HMat.hs:141-142 */
int tops_ger(tops_ctx*, const tops_buf* x, const tops_buf* y, tops_buf** out);                                /* BLAS.hs:108-110 HMat.hs:144-145 */
int tops_gemv(tops_ctx*, double alpha, const tops_buf* a, const tops_buf* x, double beta, const tops_buf* y, tops_buf** out);   /* BLAS.hs:111-116 HMat.hs:147-153 */
int tops_gemm(tops_ctx*, double alpha, const tops_buf* a, const tops_buf* b, double beta, const tops_buf* c, tops_buf** out);   /* BLAS.hs:118-123 HMat.hs:154-160 */
int tops_scale(tops_ctx*, double alpha, const tops_buf* x, tops_buf** out);                                   /* scaleB BLAS.hs:124-127 / scaleT Types.hs:70 */
int tops_add(tops_ctx*, const tops_buf* x, const tops_buf* y, tops_buf** out);                                /* addB BLAS.hs:128 — a plain add, not the reference's gemm-with-identity (BTensor.hs:113) */
int tops_index(tops_ctx*, const tops_buf* x, const int64_t* idx, double* value_out);                          /* indexB BLAS.hs:129-132 / (!) Types.hs:107-109; synchronises */
int tops_index_row(tops_ctx*, const tops_buf* a, int64_t i, tops_buf** out);                                  /* indexRowB BLAS.hs:133-136 (view, no copy) */
int tops_transp(tops_ctx*, const tops_buf* a, tops_buf** out);                                                /* transpB BLAS.hs:137-139 / transp Types.hs:71-73: reverses ALL axes */
int tops_eye(tops_ctx*, int64_t n, tops_buf** out);                                                           /* BLAS.hs:158-159 */
int tops_trace(tops_ctx*, const tops_buf* a, tops_buf** out_scalar);                                          /* traceB BLAS.hs:160-162 */
int tops_diag(tops_ctx*, int rank, const tops_buf* v, tops_buf** out);                                        /* diagB BLAS.hs:163-165 / diag Types.hs:85-88 (rank-n generalised diagonal) */
int tops_get_diag(tops_ctx*, const tops_buf* a, tops_buf** out);                                              /* getDiagB BLAS.hs:166-168 / getDiag Types.hs:89-92 */
int tops_sum(tops_ctx*, const tops_buf* x, tops_buf** out_scalar);                                            /* sumB BLAS.hs:169-171 */

/* liftB / liftT (BLAS.hs:92-96, Types.hs:56-59): n-ary elementwise map.  The reference passes a host closure
 * `Vec n a -> a`; a device cannot call it per element, so the closure is reified by applying it to symbolic
 * variables on the host side and shipping the expression as postfix bytecode (see TOPS_OP_* below).
 * A catalogue of exact programs — every single unary / binary opcode, 1/x, d*logistic'(x) as `map'` builds it, p - r*g, (a-b)^2:
 * the closures the reference's own TOps lift — is served by specialised float4 kernels (HBM-bound); anything else runs in a
 * per-thread stack interpreter — still on the device, never on the host.  tops_lift_catalogue_hits() counts the former. */
int tops_lift(tops_ctx*, const int32_t* prog, int prog_len, const float* consts, int n_consts,
              int n_in, const tops_buf* const* in, int rank, const int64_t* dims, tops_buf** out);
int64_t tops_lift_catalogue_hits(void);

enum {
    TOPS_OP_VAR = 0,    /* arg: input index      push in[arg][i]      */
    TOPS_OP_CONST = 1,  /* arg: constant index   push consts[arg]     */
    TOPS_OP_ADD = 2, TOPS_OP_SUB = 3, TOPS_OP_MUL = 4, TOPS_OP_DIV = 5,
    TOPS_OP_NEG = 6, TOPS_OP_EXP = 7, TOPS_OP_LOG = 8, TOPS_OP_RECIP = 9,
    TOPS_OP_SQRT = 10, TOPS_OP_TANH = 11, TOPS_OP_ABS = 12, TOPS_OP_SIGNUM = 13,
    TOPS_OP_MAX = 14, TOPS_OP_MIN = 15, TOPS_OP_POW = 16, TOPS_OP_LOGISTIC = 17,
    TOPS_OP_SIN = 18, TOPS_OP_COS = 19
};
/* each instruction is one int32: (opcode << 16) | arg */

/* ------------------------------------------------------------------ class Tensor beyond BLAS (Types.hs:52-109) */
/* gmul (Types.hs:60-66; semantics Data/Nested.hs:451-473):  x : ms++os, y : Reverse os ++ ns  ->  ms++ns,
 *   z[m..,n..] = sum_{o..} x[m..,o..] * y[reverse(o..),n..].   Replaces BTensor.gmulB/gmulBLAS/naiveGMul
 *   (Backend/BTensor.hs:592-716) with one tensor-core GEMM on flat storage (plus an axis permutation when |os|>=2). */
int tops_gmul(tops_ctx*, int len_m, int len_o, int len_n, const tops_buf* x, const tops_buf* y, tops_buf** out);
/* `gmul lM lO lN >>> sumRows` (TOp.hs:56-94,151-159) as one primitive and its VJP — the fusion the deferred evaluator applies when it
 * meets that composition: sumRows (gmul x y) = gmul (sumRows x) y, so the [A, ...] intermediate is never formed (the reference
 * computes A separate gemms through mapBTM, BTensor.hs:706-710, then adds them).  out: dims ms[1..] ++ ns.  VJP: ct has out's shape;
 * dx = x's shape (the cotangent broadcast over the summed axis, contracted with y), dy = y's shape. */
int tops_gmul_sum_rows(tops_ctx*, int lM, int lO, int lN, const tops_buf* x, const tops_buf* y, tops_buf** out);
int tops_gmul_sum_rows_vjp(tops_ctx*, int lM, int lO, int lN, const tops_buf* x, const tops_buf* y, const tops_buf* ct, tops_buf** dx, tops_buf** dy);
int tops_sum_t(tops_ctx*, int n, const tops_buf* const* xs, tops_buf** out);                                  /* sumT Types.hs:69 (left fold, in order) */
int tops_sum_rows(tops_ctx*, const tops_buf* x, tops_buf** out);                                              /* sumRows Types.hs:82-84 */
int tops_broadcast_rows(tops_ctx*, int64_t n, const tops_buf* row, tops_buf** out);                           /* mapRows (LS LZ) (const row): VJP of sumRows, TOp.hs:151-159 */
int tops_map_rows_softmax(tops_ctx*, const tops_buf* x, tops_buf** out);                                      /* mapRows over the leading axis of the reference's softmax TOp (NeuralNet.hs:52-59) */

/* ------------------------------------------------------------------ fused, batched hot path
 * Batched semantics: for each sample s evaluate runTOp and gradTOp' of the per-sample TOp
 *   ffLayer' >>> act      (FeedForward.hs:209-212, NeuralNet.hs:38-40)
 * with parameters fixed; outputs per sample, parameter gradients SUMMED over samples.
 *   X[B,i] W[o,i] b[o] dA[B,o]  ->  A[B,o] dX[B,i] dW[o,i] db[o]
 *   Z = X W^T + 1 b^T ; A = act(Z) ; dZ = dA ⊙ act'(Z) ; dW = dZ^T X ; db = Σ_s dZ[s,:] ; dX = dZ W        */
int tops_fflayer_fwd(tops_ctx*, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act, tops_buf** A);
/* gradTOp' of the layer: recomputes the forward unless the saved activation `A_saved` is given (Types.hs:155 recomputes) */
int tops_fflayer_grad(tops_ctx*, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act,
                      const tops_buf* dA, const tops_buf* A_saved, tops_buf** dX, tops_buf** dW, tops_buf** db);
/* forward + VJP in one call; the activation and dZ never leave the device and are produced by one GEMM epilogue */
int tops_fflayer_fwd_grad(tops_ctx*, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act,
                          const tops_buf* dA, tops_buf** A, tops_buf** dX, tops_buf** dW, tops_buf** db);
/* The same forward + VJP with the batch in HOST memory (the reference builds tensors from Haskell lists with `fromList` and reads
 * results with `toList`, Tensor.hs:187-273): X_host[B,i] and dA_host[B,o] are row-major fp32 host arrays (pinned for full PCIe
 * rate), cut into `n_chunks` row chunks (0 = default 8) whose host->device copies overlap the GEMMs of the previous chunk.
 * `grads` receives the packed device buffer [dW (o*i) || db (o)], summed over the batch; if `grads_host` is non-NULL the packed
 * gradient is also copied there and the call synchronises.  A / dX (device, may be NULL slots) receive the per-sample outputs. */
int tops_fflayer_fwd_grad_host(tops_ctx*, const float* X_host, const float* dA_host, int64_t B, const tops_buf* W, const tops_buf* b,
                               int act, int n_chunks, tops_buf** A, tops_buf** dX, tops_buf** grads, float* grads_host);
/* Data-parallel forward + VJP with the parameter-gradient all-reduce fused into the dW GEMM.  `grads_mc` is the NVLS multicast
 * alias of a packed fp32 buffer [dW (o*i) || db (o)] that every rank has bound to one multicast object (e.g. torch symmetric
 * memory's multicast_ptr).  Split-K partials are summed locally into `grads_local` (same packing); the last partial to finish a
 * region of a dW tile pushes the finished region once with multimem.red, so the NVSwitch sums it into every rank's replica while
 * the GEMM is still running; db is pushed by a tiny kernel.  Caller protocol per step: zero the symmetric replica, barrier, this
 * call, barrier.  (The reference has no parallelism; this completes the sum over samples its training fold performs serially.) */
int tops_fflayer_fwd_grad_mc(tops_ctx*, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act, const tops_buf* dA,
                             tops_buf** A, tops_buf** dX, tops_buf** grads_local, void* grads_mc);
/* Data-parallel step with the schedule owned by the library (SURVEY 8-b `tops_fflayer_step_dp`): forward, then dW and db into the
 * packed buffer `grads` = [dW (o*i) || db (o)]; the event `grads_ready` (tops_event_create) is recorded as soon as they are
 * complete and only then is the dX GEMM launched, leaving `reserve_sms` SMs free — dX does not depend on dW, so the caller's
 * all-reduce of `grads` (NCCL on a stream that waits for the event, tops_stream_wait_event) overlaps it.  Sums what the
 * reference's training fold accumulates one sample at a time (FeedForward.hs:131-148, app/Dots.hs:74-80). */
int tops_fflayer_step_dp(tops_ctx*, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act, const tops_buf* dA,
                         tops_buf** A, tops_buf** dX, tops_buf** grads, void* grads_ready, int reserve_sms);
/* ---- recorded graphs: the deferred evaluator of SURVEY 8-b (`tops_graph_begin/op/end/run`), with the API calls themselves as ops.
 * Between tops_graph_begin and tops_graph_end every call on the context is RECORDED (CUDA stream capture) instead of executed;
 * buffers allocated meanwhile (results and temporaries) come from the graph's arena (`arena_bytes`, 0 = 64 MiB) and keep their
 * addresses.  tops_graph_launch replays the whole sequence — the Category-composed forward and reverse sweep of a TOp
 * (Types.hs:135-157), an SGD step, one sample of trainNetwork's fold (FeedForward.hs:131-148) — as one cudaGraphLaunch, reading
 * its inputs from and writing its results to the same tensors every time (refresh inputs in place with tops_upload / tops_copy).
 * Not recordable: anything that reads back to the host (tops_download, tops_index, tops_sync) and the host-buffer entry point.
 * Tensors returned while recording are views into the arena: they must not outlive tops_graph_destroy. */
typedef struct tops_graph tops_graph;
int tops_graph_begin(tops_ctx*, size_t arena_bytes, tops_graph** out);
int tops_graph_end(tops_ctx*, tops_graph*);
int tops_graph_launch(tops_ctx*, tops_graph*);
int64_t tops_graph_kernel_count(const tops_graph*);   /* kernels recorded (what one launch replays) */
int tops_graph_destroy(tops_ctx*, tops_graph*);
/* dst <- src, device to device (same dtype and element count): publishes a recorded step's new state into the tensors its next
 * replay reads, e.g. the SGD-updated parameters of trainNetwork (FeedForward.hs:141-147). */
int tops_copy(tops_ctx*, tops_buf* dst, const tops_buf* src);
/* CUDA events for schedules that span streams (the handle is a cudaEvent_t). `stream` NULL = the context's current stream. */
int tops_event_create(tops_ctx*, void** ev);
int tops_event_destroy(tops_ctx*, void* ev);
int tops_stream_wait_event(tops_ctx*, void* stream, void* ev);
/* netGrad (FeedForward.hs:178-199) of a genNet-style network (FeedForward.hs:216-235) over a batch:
 *   layers l = 0..n-1 with W[l], b[l], acts[l]; loss on (A_out, Y); per-sample losses summed into loss_sum (rank 0).
 *   Outputs: A_out[B,o], loss_sum[], dX[B,i] (may be NULL to skip), dW[l], db[l]. */
int tops_mlp_fwd_grad(tops_ctx*, int n_layers, const tops_buf* const* W, const tops_buf* const* b, const int* acts, int loss,
                      const tops_buf* X, const tops_buf* Y, tops_buf** A_out, tops_buf** loss_sum, tops_buf** dX,
                      tops_buf** dW, tops_buf** db);
/* runNetwork over a batch (FeedForward.hs:123-129) */
int tops_mlp_fwd(tops_ctx*, int n_layers, const tops_buf* const* W, const tops_buf* const* b, const int* acts,
                 const tops_buf* X, tops_buf** A_out);
/* trainNetwork's parameter step  p' = p - r*g  (FeedForward.hs:141-147), for n parameter tensors at once */
int tops_sgd_step(tops_ctx*, int n, const tops_buf* const* params, const tops_buf* const* grads, double rate, tops_buf** out);

#ifdef __cplusplus
}
#endif
#endif /* TOPS_B200_H */
