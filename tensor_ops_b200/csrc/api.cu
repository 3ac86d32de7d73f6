// C ABI of libtops_b200 (include/tops_b200.h): contexts, ref-counted device tensors, the BLAS / Tensor class
// methods of tensor-ops as device operations, and the fused batched ffLayer / MLP forward+gradient paths.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <string>
#include <vector>

#include "gemm_sm100.cuh"
#include "gemm_sm100.h"
#include "kernels.h"
#include "tops_b200.h"

using namespace tops;

struct tops_ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t copy_stream = nullptr;   // H2D staging of the host-buffer entry points, overlapped with compute on `stream`
    cudaStream_t stream = nullptr;
    std::recursive_mutex mu;
    std::string last_error;
    int precision = TOPS_PREC_F16X3;
    int fused_chunk_kb = 8;
    // F16X3: k-blocks of 64 K-elements (12 MMAs) per TMEM chunk.  The forward GEMM's error is amplified ~4x by act' at a saturating
    // init, so it drains every 2 k-blocks (6e-7); the gradient GEMMs every 4 (1.1e-6, 7 % faster).  TOPS_F16X3_CHUNK / _FWD_CHUNK override.
    int f16x3_chunk_kb = 4;
    int f16x3_fwd_chunk_kb = 2;
    bool skinny = true;           // products with a dimension <= 16 take the streaming fp32 kernels; TOPS_SKINNY=0: they stay on the tcgen05 path (padded fp16 planes)
    int f16x3_fwd_head_kb = 4;    // first two chunks of every tile of the fused forward GEMM (lookahead for its epilogue); TOPS_F16X3_FWD_HEAD
    struct SplitEntry { const void* src; int64_t rows, cols; void* hi; void* lo; long long ld; float* scale2; };
    struct SplitScope* split_scope = nullptr;   // fp16 pairs already made inside the current API call
    struct tops_graph* capturing = nullptr;     // non-NULL between tops_graph_begin and tops_graph_end: allocations come from its arena
    int64_t launches = 0;
    // lifetime: one reference for the handle returned by tops_init + one per live tops_buf; the struct is deleted by whoever drops the
    // last one, so a buffer finalizer that runs after tops_shutdown (a host GC) still finds its context
    std::atomic<int> refs{1};
    bool shut = false;
    unsigned int* wd_host = nullptr;
    unsigned int* wd_dev = nullptr;
    float* scratch = nullptr;   // small persistent workspace for reductions (1 MiB)
    // optional per-kernel timing (tops_profile_*): CUDA events recorded on ctx->stream around tagged launches
    struct ProfRec { const char* tag; cudaEvent_t e0, e1; double flops, bytes; };
    bool profiling = false;
    std::vector<ProfRec> prof;
};

// A recorded sequence of API calls (tops_graph_begin .. tops_graph_end) replayed as ONE CUDA graph launch.  Every device buffer
// allocated while recording — results and temporaries alike — lives in the graph's arena at a fixed address, so a replay writes
// its results to the same tensors the recording call returned.
struct tops_graph {
    tops_ctx* ctx = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    char* arena = nullptr;
    size_t arena_bytes = 0, used = 0;
    int64_t launches_recorded = 0;
};

struct tops_buf {
    tops_ctx* ctx;
    std::atomic<int> refs;
    int dtype;
    int rank;
    int64_t dims[TOPS_MAX_RANK];   // logical dims
    int64_t numel;
    void* data;
    bool owns;
    bool tr;           // rank-2 only: storage is the row-major TRANSPOSE of the logical matrix (O(1) transp, like hmatrix `tr`)
    tops_buf* parent;  // kept alive while this view lives
};

namespace {

constexpr size_t kScratchBytes = 4u << 20;

int set_err(tops_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->last_error = buf;
    return code;
}

// serialise calls on the context and make its device current (a process may hold contexts on several devices)
// ... and restore the caller's device afterwards (GC finalizers of a host runtime call in here from arbitrary threads)
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};
#define LOCK(ctx) std::lock_guard<std::recursive_mutex> lock_((ctx)->mu); DeviceGuard dev_guard_((ctx)->device)
#define CHECK_CTX(ctx) do { if (!(ctx)) return TOPS_ERR_INVALID; } while (0)
#define CUDA_TRY(ctx, expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return set_err(ctx, TOPS_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); } while (0)
#define TRY(expr) do { int r_ = (expr); if (r_ != TOPS_OK) return r_; } while (0)

size_t esize(int dtype) { return dtype == TOPS_BF16 ? 2 : 4; }

k::LaunchCtx lc_of(tops_ctx* ctx) { return k::LaunchCtx{ctx->stream, ctx->num_sms, &ctx->launches}; }

int check_launch(tops_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_err(ctx, TOPS_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    return TOPS_OK;
}

// RAII: brackets the launches issued in its scope with two events when profiling is on (no-op otherwise)
struct ProfScope {
    tops_ctx* ctx; bool on = false; tops_ctx::ProfRec r{};
    ProfScope(tops_ctx* c, const char* tag, double flops, double bytes) : ctx(c) {
        if (!c->profiling) return;
        if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) { cudaGetLastError(); return; }
        r.tag = tag ? tag : "untagged"; r.flops = flops; r.bytes = bytes;
        cudaEventRecord(r.e0, c->stream);
        on = true;
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(r.e1, ctx->stream);
        ctx->prof.push_back(r);
    }
};

int new_buf(tops_ctx* ctx, int dtype, int rank, const int64_t* dims, void* data, bool owns, tops_buf* parent, tops_buf** out) {
    if (rank < 0 || rank > TOPS_MAX_RANK) return set_err(ctx, TOPS_ERR_INVALID, "rank %d out of range", rank);
    tops_buf* b = new tops_buf();
    b->ctx = ctx; b->refs = 1; b->dtype = dtype; b->rank = rank; b->numel = 1;
    for (int i = 0; i < rank; ++i) {
        if (dims[i] < 0) { delete b; return set_err(ctx, TOPS_ERR_INVALID, "negative dimension"); }
        b->dims[i] = dims[i]; b->numel *= dims[i];
    }
    b->data = data; b->owns = owns; b->tr = false; b->parent = parent;
    if (parent) parent->refs.fetch_add(1);
    ctx->refs.fetch_add(1);
    *out = b;
    return TOPS_OK;
}

int arena_alloc(tops_ctx* ctx, size_t bytes, void** p) {
    tops_graph* g = ctx->capturing;
    const size_t need = (bytes + 255) & ~(size_t)255;
    if (g->used + need > g->arena_bytes)
        return set_err(ctx, TOPS_ERR_OOM, "graph arena exhausted: %zu bytes used of %zu, %zu more requested (pass a larger arena to tops_graph_begin)", g->used, g->arena_bytes, need);
    *p = g->arena + g->used;
    g->used += need;
    return TOPS_OK;
}

int alloc_buf(tops_ctx* ctx, int dtype, int rank, const int64_t* dims, tops_buf** out) {
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dims[i];
    void* p = nullptr;
    size_t bytes = (size_t)(n > 0 ? n : 1) * esize(dtype);
    if (ctx->capturing) {   // recording a graph: bump-allocate from its arena (fixed addresses across replays; freed with the graph)
        TRY(arena_alloc(ctx, bytes, &p));
        return new_buf(ctx, dtype, rank, dims, p, false, nullptr, out);
    }
    cudaError_t e = cudaMallocAsync(&p, bytes, ctx->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return set_err(ctx, e == cudaErrorMemoryAllocation ? TOPS_ERR_OOM : TOPS_ERR_CUDA, "device allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    int r = new_buf(ctx, dtype, rank, dims, p, true, nullptr, out);
    if (r != TOPS_OK) cudaFreeAsync(p, ctx->stream);
    return r;
}

// raw workspace for one API call: stream-ordered normally, from the arena while a graph is being recorded
int ws_alloc(tops_ctx* ctx, size_t bytes, void** p) {
    if (ctx->capturing) return arena_alloc(ctx, bytes, p);
    cudaError_t e = cudaMallocAsync(p, bytes ? bytes : 1, ctx->stream);
    if (e != cudaSuccess) { cudaGetLastError(); return set_err(ctx, TOPS_ERR_OOM, "workspace allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); }
    return TOPS_OK;
}
void ws_free(tops_ctx* ctx, void* p) { if (p && !ctx->capturing) cudaFreeAsync(p, ctx->stream); }

void release_buf(tops_buf* b) {
    while (b) {
        if (b->refs.fetch_sub(1) != 1) return;
        tops_buf* parent = b->parent;
        tops_ctx* ctx = b->ctx;
        if (b->owns && b->data) { if (ctx->shut) cudaFree(b->data); else cudaFreeAsync(b->data, ctx->stream); }
        delete b;
        if (ctx->refs.fetch_sub(1) == 1) delete ctx;      // only possible after tops_shutdown dropped the handle's reference
        b = parent;
    }
}

struct Tmp {   // RAII holder for temporaries created inside one API call
    std::vector<tops_buf*> v;
    ~Tmp() { for (auto* b : v) release_buf(b); }
    tops_buf* keep(tops_buf* b) { v.push_back(b); return b; }
};

bool same_shape(const tops_buf* a, const tops_buf* b) {
    if (a->rank != b->rank) return false;
    for (int i = 0; i < a->rank; ++i) if (a->dims[i] != b->dims[i]) return false;
    return true;
}

// Output protocol: *out == NULL -> allocate; otherwise validate and write in place.
int prep_out(tops_ctx* ctx, tops_buf** out, int dtype, int rank, const int64_t* dims, bool* fresh = nullptr) {
    if (!out) return set_err(ctx, TOPS_ERR_INVALID, "NULL output slot");
    if (fresh) *fresh = (*out == nullptr);
    if (*out == nullptr) return alloc_buf(ctx, dtype, rank, dims, out);
    tops_buf* o = *out;
    if (o->tr) return set_err(ctx, TOPS_ERR_INVALID, "pre-allocated output must not be a transposed view");
    if (o->dtype != dtype) return set_err(ctx, TOPS_ERR_SHAPE, "pre-allocated output has dtype %d, expected %d", o->dtype, dtype);
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dims[i];
    if (o->numel != n) return set_err(ctx, TOPS_ERR_SHAPE, "pre-allocated output has %lld elements, expected %lld", (long long)o->numel, (long long)n);
    return TOPS_OK;
}

// contiguous (non-transposed) version of x; returns x itself (retained) when already contiguous
int contig(tops_ctx* ctx, const tops_buf* x, Tmp& tmp, const tops_buf** out) {
    if (!x->tr) { *out = x; return TOPS_OK; }
    if (x->dtype != TOPS_F32) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "transposed bf16 view cannot be materialised");
    tops_buf* t = nullptr;
    TRY(alloc_buf(ctx, x->dtype, x->rank, x->dims, &t));
    tmp.keep(t);
    // storage is [dims1, dims0] row-major; logical [dims0, dims1] = transpose of storage
    k::transpose2d(lc_of(ctx), (const float*)x->data, (float*)t->data, x->dims[1], x->dims[0]);
    *out = t;
    return check_launch(ctx, "transpose");
}

int need_f32(tops_ctx* ctx, const tops_buf* x, const char* what) {
    if (!x) return set_err(ctx, TOPS_ERR_INVALID, "%s: NULL tensor", what);
    if (x->dtype != TOPS_F32) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "%s: only fp32 tensors are supported here", what);
    return TOPS_OK;
}

// ---------------------------------------------------------------------------------------------- GEMM dispatch
// F16X3: an fp32 [rows, cols] operand as an fp16 pair of x * 2^k with one k for the whole tensor (absmax + split passes).  Within
// one API call (SplitScope) the pair of a given buffer is made once and reused by every GEMM that reads it.
}  // namespace
struct SplitScope {
    tops_ctx* ctx; std::vector<tops_ctx::SplitEntry> entries; Tmp keep; bool owner;
    // GEMM outputs whose max|.| the producing epilogue has already reduced (GemmCall::absmax_out): their split skips the absmax pass
    struct KnownMax { const void* ptr; int64_t numel; unsigned* bits; };
    std::vector<KnownMax> known_max;
    bool track_max = false;          // set by calls whose GEMM outputs are operands of later GEMMs of the same call (the MLP)
    explicit SplitScope(tops_ctx* c) : ctx(c), owner(c->split_scope == nullptr) { if (owner) c->split_scope = this; }
    ~SplitScope() { if (owner) ctx->split_scope = nullptr; }
};
namespace {
constexpr int kF16X3Fallback = -1000;   // the operands cannot be expressed: the caller takes the TF32_BF16X2 route instead

int split_operand_f16(tops_ctx* ctx, const void* src, int64_t rows, int64_t cols, Tmp& tmp, tops_ctx::SplitEntry* out) {
    SplitScope* scope = ctx->split_scope;
    if (scope)
        for (auto& e : scope->entries)
            if (e.src == src && e.rows == rows && e.cols == cols) { *out = e; return TOPS_OK; }
    Tmp& holder = scope ? scope->keep : tmp;   // cached pairs live until the API call ends
    const int64_t ld = (cols + 7) / 8 * 8;   // 16-byte aligned fp16 rows for every shape
    int64_t pd[1] = {rows * ld}, sd[1] = {4};
    tops_buf *hi = nullptr, *lo = nullptr, *sc = nullptr;
    TRY(alloc_buf(ctx, TOPS_BF16, 1, pd, &hi)); holder.keep(hi);
    TRY(alloc_buf(ctx, TOPS_BF16, 1, pd, &lo)); holder.keep(lo);
    TRY(alloc_buf(ctx, TOPS_F32, 1, sd, &sc)); holder.keep(sc);   // {scale, 1/scale, absmax bits, -}
    unsigned* mx = nullptr;
    if (scope)
        for (auto& km : scope->known_max)
            if (km.ptr == src && km.numel == rows * cols) mx = km.bits;
    if (mx == nullptr) {
        mx = reinterpret_cast<unsigned*>(sc->data) + 2;
        CUDA_TRY(ctx, cudaMemsetAsync(mx, 0, 4, ctx->stream));
        k::absmax_bits(lc_of(ctx), (const float*)src, rows * cols, mx);
    }
    k::split_f16_tensor_2d(lc_of(ctx), (const float*)src, rows, cols, ld, mx, hi->data, lo->data, (float*)sc->data);
    TRY(check_launch(ctx, "split_f16"));
    *out = tops_ctx::SplitEntry{src, rows, cols, hi->data, lo->data, ld, (float*)sc->data};
    if (scope) scope->entries.push_back(*out);
    return TOPS_OK;
}

int run_gemm(tops_ctx* ctx, GemmCall c);

// Inside a call that tracks them (SplitScope::track_max): give the GEMM a zeroed slot where its epilogue leaves max|out0|, so that the
// fp16-pair split of out0 later in the same call needs no absmax pass.  *done is set by the launcher when the variant produced it.
int want_out_max(tops_ctx* ctx, GemmCall& c, int* done) {
    SplitScope* scope = ctx->split_scope;
    if (!(scope && scope->track_max && (c.epi == EPI_BIAS_ACT || c.epi == EPI_MUL_DACT) && c.ld_out0 == c.N && c.out0_mc == nullptr)) return TOPS_OK;
    int64_t sd[1] = {1};
    tops_buf* mxb = nullptr;
    TRY(alloc_buf(ctx, TOPS_F32, 1, sd, &mxb)); scope->keep.keep(mxb);
    CUDA_TRY(ctx, cudaMemsetAsync(mxb->data, 0, 4, ctx->stream));
    c.absmax_out = reinterpret_cast<unsigned*>(mxb->data);
    c.absmax_done = done;
    return TOPS_OK;
}

// fp32 operands under TOPS_PREC_F16X3: split both (cached per call) and run the fp16-pair kernel
int run_gemm_f16x3(tops_ctx* ctx, const GemmCall& c0) {
    if (c0.dtype != 0 || c0.io_bf16 || c0.M <= 0 || c0.N <= 0 || c0.K <= 0) return kF16X3Fallback;
    const int64_t a_rows = c0.major_a == MAJOR_K ? c0.M : c0.K, a_cols = c0.major_a == MAJOR_K ? c0.K : c0.M;
    const int64_t b_rows = c0.major_b == MAJOR_K ? c0.N : c0.K, b_cols = c0.major_b == MAJOR_K ? c0.K : c0.N;
    if (c0.lda != a_cols || c0.ldb != b_cols) return kF16X3Fallback;   // strided views: not produced by this library
    Tmp tmp;
    tops_ctx::SplitEntry a, b;
    TRY(split_operand_f16(ctx, c0.A, a_rows, a_cols, tmp, &a));
    TRY(split_operand_f16(ctx, c0.B, b_rows, b_cols, tmp, &b));
    GemmCall c = c0;
    c.dtype = 2; c.passes = 0;
    c.A = a.hi; c.A2 = a.lo; c.lda = a.ld;
    c.B = b.hi; c.B2 = b.lo; c.ldb = b.ld;
    c.B16 = c.Blo16 = nullptr;
    c.acc_scale_ptr = a.scale2 + 1; c.acc_scale_ptr2 = b.scale2 + 1;   // 1 / sA and 1 / sB (powers of two), multiplied in the epilogue
    SplitScope* scope = ctx->split_scope;
    int max_done = 0;
    TRY(want_out_max(ctx, c, &max_done));
    if (c.chunk_kb <= 0) {
        c.chunk_kb = c0.aux0 != nullptr ? ctx->f16x3_fwd_chunk_kb : ctx->f16x3_chunk_kb;
        if (c0.aux0 != nullptr) c.chunk_head_kb = ctx->f16x3_fwd_head_kb;
    }
    const int r = run_gemm(ctx, c);
    if (r == TOPS_OK && max_done) scope->known_max.push_back({c.out0, (int64_t)c.M * c.N, c.absmax_out});
    return r;
}

int run_gemm(tops_ctx* ctx, GemmCall c) {
    if (c.colsum_fused) *c.colsum_fused = 0;
    if (c.M <= 0 || c.N <= 0) return TOPS_OK;
    // One tiny dimension (an MLP's output layer): streaming fp32 CUDA-core kernels — HBM-bound work the tensor-core tiles and the
    // operand splits would only slow down.  FP32_SIMT keeps its single reference kernel.
    if (ctx->skinny && ctx->precision != TOPS_PREC_FP32_SIMT && k::gemm_skinny_kind(c) != 0) {
        ProfScope prof_(ctx, c.tag ? c.tag : "gemm", 2.0 * c.M * c.N * (double)c.K,
                        4.0 * ((double)c.M * c.K + (double)c.N * c.K + (double)c.M * c.N * ((c.aux0 ? 1 : 0) + 1)));
        if (c.epi == EPI_ATOMIC && !c.accumulate) CUDA_TRY(ctx, cudaMemsetAsync(c.out0, 0, sizeof(float) * (size_t)c.M * (size_t)c.ld_out0, ctx->stream));
        int max_done = 0;
        TRY(want_out_max(ctx, c, &max_done));
        const int r = k::gemm_skinny(lc_of(ctx), c, c.colsum_fused, &max_done);
        if (r < 0) return set_err(ctx, TOPS_ERR_CUDA, "skinny gemm launch failed (%d)", -r);
        if (max_done) ctx->split_scope->known_max.push_back({c.out0, (int64_t)c.M * c.N, c.absmax_out});
        return TOPS_OK;
    }
    // F16X3 with fp32 operands: split them into fp16 pairs (comes back here with dtype 2).  Small products (< 2 GFLOP) are not worth
    // the split passes: they take the in-kernel TF32 + bf16-correction route below when TMA can describe them as they are.
    const bool f16x3 = c.dtype == 0 && ctx->precision == TOPS_PREC_F16X3;
    const bool small = 2.0 * c.M * c.N * (double)c.K < 2147483648.0;
    if (f16x3 && !small) {
        const int r = run_gemm_f16x3(ctx, c);
        if (r != kF16X3Fallback) return r;
    }
    const double es_in = c.dtype == 1 ? 2.0 : 4.0, es_out = c.io_bf16 ? 2.0 : 4.0;   // fp16 pairs: 2 x 2 bytes per element
    ProfScope prof_(ctx, c.tag ? c.tag : "gemm", 2.0 * c.M * c.N * (double)(c.K > 0 ? c.K : 0),
                    es_in * ((double)c.M * c.K + (double)c.N * c.K) + (c.epi == EPI_ATOMIC ? 4.0 : es_out) * (double)c.M * c.N * ((c.out1 ? 1 : 0) + (c.aux0 ? 1 : 0) + 1));
    if (c.epi == EPI_ATOMIC && !c.accumulate) {
        CUDA_TRY(ctx, cudaMemsetAsync(c.out0, 0, sizeof(float) * (size_t)c.M * (size_t)c.ld_out0, ctx->stream));
    }
    if (c.K <= 0) {   // empty contraction: result is the epilogue applied to zeros; only plain store/atomic make sense
        if (c.epi == EPI_STORE && !c.aux0) CUDA_TRY(ctx, cudaMemsetAsync(c.out0, 0, esize(c.io_bf16 ? TOPS_BF16 : TOPS_F32) * (size_t)c.M * (size_t)c.ld_out0, ctx->stream));
        if (c.epi == EPI_STORE || c.epi == EPI_ATOMIC) return TOPS_OK;
        return set_err(ctx, TOPS_ERR_SHAPE, "gemm: empty contraction dimension");
    }
    const bool want_umma = c.dtype != 0 || ctx->precision != TOPS_PREC_FP32_SIMT;
    if (want_umma) {
        c.passes = c.dtype != 0 ? 1 : ctx->precision == TOPS_PREC_TF32X3 ? 3 : (ctx->precision == TOPS_PREC_TF32_BF16X2 || ctx->precision == TOPS_PREC_F16X3) ? 2 : 1;
        char err[256];
        int r = gemm_umma_launch(c, ctx->stream, ctx->wd_dev, ctx->num_sms, err, sizeof err);
        if (r == 0) { ++ctx->launches; return TOPS_OK; }
        if (r > 0) return set_err(ctx, TOPS_ERR_CUDA, "%s", err);
        if (f16x3 && small && c.K > 0) {   // rows TMA cannot describe (e.g. 10 columns): the split pads them, so the tensor cores still apply
            if (c.epi == EPI_ATOMIC) c.accumulate = 1;   // already zeroed above
            const int r2 = run_gemm_f16x3(ctx, c);
            if (r2 != kF16X3Fallback) return r2;
        }
        if (c.dtype != 0) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "16-bit gemm needs 16-byte aligned operands with strides that are multiples of 8 elements (%s)", err);
    }
    if (c.out0_mc) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "the fused all-reduce needs the tcgen05 GEMM: 16-byte aligned operands and rows");
    int r = k::gemm_simt(lc_of(ctx), c);
    if (r != 0) return set_err(ctx, TOPS_ERR_CUDA, "simt gemm launch failed (%d)", r);
    return TOPS_OK;
}

// A plain-store GEMM with few output tiles and a long contraction (e.g. the dy VJP of a rank-3 gmul: 64x64 outputs, K = 4096)
// would run on a handful of SMs: switch it to split-K partials accumulated with fp32 reds (run_gemm zeroes the output).
void split_k_if_skinny(tops_ctx* ctx, GemmCall& g) {
    if (g.epi != EPI_STORE || g.aux0 != nullptr || g.io_bf16) return;
    const long long tiles = ((g.M + 127) / 128) * (long long)((g.N + 255) / 256);
    if (tiles * 4 <= ctx->num_sms && g.K >= 1024) { g.epi = EPI_ATOMIC; g.split_k = 0; }
}

// describe a rank-2 fp32/bf16 matrix (possibly a transposed view) as a GEMM operand whose reduction runs over `k_axis`
// of its LOGICAL shape.  Returns pointer/ld/major for the engine's  sum_k P(mn, k)  convention.
void as_operand(const tops_buf* m, int k_axis, const void** ptr, long long* ld, int* major) {
    // logical [d0, d1]; stored row-major as [d0,d1] (tr=0) or [d1,d0] (tr=1)
    const bool k_is_contiguous = (k_axis == 1) != m->tr;
    *ptr = m->data;
    *ld = m->tr ? m->dims[0] : m->dims[1];
    *major = k_is_contiguous ? MAJOR_K : MAJOR_MN;
}

}  // namespace

// ================================================================================================ lifecycle
extern "C" int tops_init(int device, tops_ctx** out) {
    if (!out) return TOPS_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) { cudaGetLastError(); return TOPS_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return TOPS_ERR_NO_DEVICE;
    if (prop.major != 10) return TOPS_ERR_NO_DEVICE;   // sm_100a kernels only; there is no fallback path
    if (cudaSetDevice(device) != cudaSuccess) return TOPS_ERR_CUDA;
    tops_ctx* ctx = new tops_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return TOPS_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { cudaStreamDestroy(ctx->own_stream); delete ctx; return TOPS_ERR_CUDA; }
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    if (cudaHostAlloc((void**)&ctx->wd_host, 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer((void**)&ctx->wd_dev, ctx->wd_host, 0) != cudaSuccess ||
        cudaMalloc((void**)&ctx->scratch, kScratchBytes) != cudaSuccess) {
        cudaGetLastError();
        delete ctx;
        return TOPS_ERR_CUDA;
    }
    memset(ctx->wd_host, 0, 64);
    if (const char* e = getenv("TOPS_F16X3_CHUNK")) { const int v = atoi(e); if (v >= 1 && v <= 64) ctx->f16x3_chunk_kb = v; }
    if (const char* e = getenv("TOPS_F16X3_FWD_CHUNK")) { const int v = atoi(e); if (v >= 1 && v <= 64) ctx->f16x3_fwd_chunk_kb = v; }
    if (const char* e = getenv("TOPS_SKINNY")) ctx->skinny = atoi(e) != 0;
    if (const char* e = getenv("TOPS_F16X3_FWD_HEAD")) { const int v = atoi(e); if (v >= 0 && v <= 64) ctx->f16x3_fwd_head_kb = v; }
    *out = ctx;
    return TOPS_OK;
}

extern "C" int tops_shutdown(tops_ctx* ctx) {
    CHECK_CTX(ctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->wd_host) cudaFreeHost(ctx->wd_host);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    ctx->scratch = nullptr; ctx->wd_host = nullptr; ctx->copy_stream = ctx->own_stream = ctx->stream = nullptr;
    ctx->shut = true;                                      // buffers released from now on free synchronously and make no launches
    if (ctx->refs.fetch_sub(1) == 1) delete ctx;
    return TOPS_OK;
}

extern "C" const char* tops_last_error(tops_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "NULL context"; }

extern "C" int tops_sync(tops_ctx* ctx) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (ctx->capturing) return set_err(ctx, TOPS_ERR_INVALID, "a graph is being recorded: calls that read results back to the host (download, index, sync) are not recordable");
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess)
        return set_err(ctx, TOPS_ERR_CUDA, "stream sync failed: %s (gemm watchdog code=0x%x cta=%u)", cudaGetErrorString(e), ctx->wd_host[0], ctx->wd_host[1]);
    return TOPS_OK;
}
extern "C" int tops_set_stream(tops_ctx* ctx, void* s) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (ctx->capturing) return set_err(ctx, TOPS_ERR_INVALID, "tops_set_stream while a graph is being recorded");
    // work queued on the previous stream may still be using buffers that the next stream's allocations would recycle
    if (ctx->stream != (s ? (cudaStream_t)s : ctx->own_stream)) cudaStreamSynchronize(ctx->stream);
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return TOPS_OK;
}
extern "C" int tops_set_precision(tops_ctx* ctx, int p) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (p < 0 || p > 4) return set_err(ctx, TOPS_ERR_INVALID, "unknown precision %d", p);
    ctx->precision = p;
    return TOPS_OK;
}
extern "C" int tops_get_precision(tops_ctx* ctx) { return ctx ? ctx->precision : -1; }
extern "C" int tops_profile_enable(tops_ctx* ctx, int on) {
    CHECK_CTX(ctx); LOCK(ctx);
    ctx->profiling = on != 0;
    return TOPS_OK;
}
// Synchronises the stream, aggregates the recorded event pairs per tag and clears them.  JSON:
//   {"<tag>": {"launches": n, "ms": total, "flops": total, "bytes": total}, ...}
extern "C" int tops_profile_summary(tops_ctx* ctx, char* out, size_t cap) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!out || cap < 3) return set_err(ctx, TOPS_ERR_INVALID, "tops_profile_summary: NULL/short output buffer");
    TRY(tops_sync(ctx));
    struct Agg { std::string tag; int n; double ms, flops, bytes; };
    std::vector<Agg> aggs;
    for (auto& r : ctx->prof) {
        float ms = 0.f;
        cudaEventSynchronize(r.e1);
        cudaEventElapsedTime(&ms, r.e0, r.e1);
        cudaEventDestroy(r.e0); cudaEventDestroy(r.e1);
        Agg* a = nullptr;
        for (auto& x : aggs) if (x.tag == r.tag) a = &x;
        if (!a) { aggs.push_back(Agg{r.tag, 0, 0, 0, 0}); a = &aggs.back(); }
        a->n += 1; a->ms += ms; a->flops += r.flops; a->bytes += r.bytes;
    }
    ctx->prof.clear();
    std::string js = "{";
    for (size_t i = 0; i < aggs.size(); ++i) {
        char buf[256];
        snprintf(buf, sizeof buf, "%s\"%s\": {\"launches\": %d, \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}", i ? ", " : "", aggs[i].tag.c_str(), aggs[i].n, aggs[i].ms, aggs[i].flops, aggs[i].bytes);
        js += buf;
    }
    js += "}";
    if (js.size() + 1 > cap) return set_err(ctx, TOPS_ERR_INVALID, "tops_profile_summary: buffer of %zu bytes too small (%zu needed)", cap, js.size() + 1);
    memcpy(out, js.c_str(), js.size() + 1);
    return TOPS_OK;
}
extern "C" int64_t tops_launch_count(tops_ctx* ctx) { return ctx ? ctx->launches : -1; }
extern "C" int64_t tops_lift_catalogue_hits(void) { return k::g_lift_catalogue_hits; }
extern "C" int tops_device_sm_count(tops_ctx* ctx) { return ctx ? ctx->num_sms : -1; }

// ================================================================================================ storage
extern "C" int tops_buf_alloc(tops_ctx* ctx, int dtype, int rank, const int64_t* dims, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!out || (rank > 0 && !dims)) return set_err(ctx, TOPS_ERR_INVALID, "tops_buf_alloc: NULL argument");
    if (dtype != TOPS_F32 && dtype != TOPS_BF16) return set_err(ctx, TOPS_ERR_INVALID, "unknown dtype %d", dtype);
    *out = nullptr;
    return alloc_buf(ctx, dtype, rank, dims, out);
}
extern "C" int tops_buf_wrap(tops_ctx* ctx, void* p, int dtype, int rank, const int64_t* dims, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!out || !p) return set_err(ctx, TOPS_ERR_INVALID, "tops_buf_wrap: NULL argument");
    return new_buf(ctx, dtype, rank, dims, p, false, nullptr, out);
}
extern "C" int tops_buf_view(tops_ctx* ctx, tops_buf* parent, int64_t offset, int rank, const int64_t* dims, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!parent || !out) return set_err(ctx, TOPS_ERR_INVALID, "tops_buf_view: NULL argument");
    if (parent->tr) return set_err(ctx, TOPS_ERR_INVALID, "cannot take a flat view of a transposed view");
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dims[i];
    if (offset < 0 || offset + n > parent->numel) return set_err(ctx, TOPS_ERR_SHAPE, "view [%lld, %lld) exceeds parent of %lld elements", (long long)offset, (long long)(offset + n), (long long)parent->numel);
    return new_buf(ctx, parent->dtype, rank, dims, (char*)parent->data + offset * esize(parent->dtype), false, parent, out);
}
extern "C" int tops_buf_retain(tops_buf* b) { if (!b) return TOPS_ERR_INVALID; b->refs.fetch_add(1); return TOPS_OK; }
extern "C" int tops_buf_release(tops_buf* b) {
    if (!b) return TOPS_ERR_INVALID;
    tops_ctx* ctx = b->ctx;
    ctx->refs.fetch_add(1);                 // keep the context (and the mutex we hold) alive across a release that drops its last buffer
    {
        LOCK(ctx);
        release_buf(b);
    }
    if (ctx->refs.fetch_sub(1) == 1) delete ctx;
    return TOPS_OK;
}
extern "C" int tops_buf_rank(const tops_buf* b) { return b ? b->rank : -1; }
extern "C" int tops_buf_dims(const tops_buf* b, int64_t* d) { if (!b || !d) return TOPS_ERR_INVALID; for (int i = 0; i < b->rank; ++i) d[i] = b->dims[i]; return TOPS_OK; }
extern "C" int tops_buf_dtype(const tops_buf* b) { return b ? b->dtype : -1; }
extern "C" int64_t tops_buf_numel(const tops_buf* b) { return b ? b->numel : -1; }
extern "C" void* tops_buf_data(const tops_buf* b) { return b ? b->data : nullptr; }

extern "C" int tops_host_alloc(size_t bytes, int write_combined, void** out) {
    if (!out) return TOPS_ERR_INVALID;
    *out = nullptr;
    cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, write_combined ? cudaHostAllocWriteCombined : cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return e == cudaErrorMemoryAllocation ? TOPS_ERR_OOM : TOPS_ERR_CUDA; }
    return TOPS_OK;
}
extern "C" int tops_host_free(void* p) {
    if (!p) return TOPS_ERR_INVALID;
    return cudaFreeHost(p) == cudaSuccess ? TOPS_OK : TOPS_ERR_CUDA;
}
extern "C" int tops_upload(tops_ctx* ctx, tops_buf* dst, const void* host, size_t bytes) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!dst || (!host && bytes)) return set_err(ctx, TOPS_ERR_INVALID, "tops_upload: NULL argument");
    if (dst->tr) return set_err(ctx, TOPS_ERR_INVALID, "cannot upload into a transposed view");
    if (bytes != (size_t)dst->numel * esize(dst->dtype)) return set_err(ctx, TOPS_ERR_SHAPE, "upload of %zu bytes into a tensor of %zu bytes", bytes, (size_t)dst->numel * esize(dst->dtype));
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(dst->data, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return TOPS_OK;
}
extern "C" int tops_download(tops_ctx* ctx, const tops_buf* src, void* host, size_t bytes) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!src || (!host && bytes)) return set_err(ctx, TOPS_ERR_INVALID, "tops_download: NULL argument");
    if (ctx->capturing) return set_err(ctx, TOPS_ERR_INVALID, "tops_download is not recordable (it reads back to the host)");
    Tmp tmp; const tops_buf* s;
    TRY(contig(ctx, src, tmp, &s));
    if (bytes != (size_t)s->numel * esize(s->dtype)) return set_err(ctx, TOPS_ERR_SHAPE, "download of %zu bytes from a tensor of %zu bytes", bytes, (size_t)s->numel * esize(s->dtype));
    if (bytes) CUDA_TRY(ctx, cudaMemcpyAsync(host, s->data, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return tops_sync(ctx);
}
extern "C" int tops_fill(tops_ctx* ctx, tops_buf* dst, double v) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!dst) return set_err(ctx, TOPS_ERR_INVALID, "tops_fill: NULL tensor");
    if (dst->dtype == TOPS_BF16) k::fill_bf16(lc_of(ctx), dst->data, dst->numel, (float)v);
    else k::fill(lc_of(ctx), (float*)dst->data, dst->numel, (float)v);
    return check_launch(ctx, "fill");
}
extern "C" int tops_rand_normal(tops_ctx* ctx, tops_buf* dst, double mean, double sd, uint64_t seed) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, dst, "tops_rand_normal"));
    k::rand_normal(lc_of(ctx), (float*)dst->data, dst->numel, (float)mean, (float)sd, seed);
    return check_launch(ctx, "rand_normal");
}
extern "C" int tops_rand_uniform(tops_ctx* ctx, tops_buf* dst, double lo, double hi, uint64_t seed) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, dst, "tops_rand_uniform"));
    k::rand_uniform(lc_of(ctx), (float*)dst->data, dst->numel, (float)lo, (float)hi, seed);
    return check_launch(ctx, "rand_uniform");
}
extern "C" int tops_cast(tops_ctx* ctx, const tops_buf* x, int dtype, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!x) return set_err(ctx, TOPS_ERR_INVALID, "tops_cast: NULL tensor");
    Tmp tmp; const tops_buf* s;
    TRY(contig(ctx, x, tmp, &s));
    TRY(prep_out(ctx, out, dtype, s->rank, s->dims));
    if (s->dtype == dtype) CUDA_TRY(ctx, cudaMemcpyAsync((*out)->data, s->data, (size_t)s->numel * esize(dtype), cudaMemcpyDeviceToDevice, ctx->stream));
    else if (dtype == TOPS_BF16) k::cast_f32_bf16(lc_of(ctx), (const float*)s->data, (*out)->data, s->numel);
    else k::cast_bf16_f32(lc_of(ctx), s->data, (float*)(*out)->data, s->numel);
    return check_launch(ctx, "cast");
}

// ================================================================================================ class BLAS
extern "C" int tops_axpy(tops_ctx* ctx, double alpha, const tops_buf* x, const tops_buf* y, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, x, "tops_axpy"));
    if (y) { TRY(need_f32(ctx, y, "tops_axpy")); if (y->numel != x->numel) return set_err(ctx, TOPS_ERR_SHAPE, "axpy: x has %lld elements, y has %lld", (long long)x->numel, (long long)y->numel); }
    Tmp tmp; const tops_buf *xs, *ys = nullptr;
    TRY(contig(ctx, x, tmp, &xs));
    if (y) TRY(contig(ctx, y, tmp, &ys));
    TRY(prep_out(ctx, out, TOPS_F32, xs->rank, xs->dims));
    k::axpy(lc_of(ctx), (float)alpha, (const float*)xs->data, ys ? (const float*)ys->data : nullptr, (float*)(*out)->data, xs->numel);
    return check_launch(ctx, "axpy");
}
extern "C" int tops_scale(tops_ctx* ctx, double alpha, const tops_buf* x, tops_buf** out) { return tops_axpy(ctx, alpha, x, nullptr, out); }
extern "C" int tops_add(tops_ctx* ctx, const tops_buf* x, const tops_buf* y, tops_buf** out) {
    CHECK_CTX(ctx);
    if (!y) return set_err(ctx, TOPS_ERR_INVALID, "tops_add: NULL tensor");
    return tops_axpy(ctx, 1.0, x, y, out);
}
extern "C" int tops_dot(tops_ctx* ctx, const tops_buf* x, const tops_buf* y, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, x, "tops_dot")); TRY(need_f32(ctx, y, "tops_dot"));
    if (x->numel != y->numel) return set_err(ctx, TOPS_ERR_SHAPE, "dot: %lld vs %lld elements", (long long)x->numel, (long long)y->numel);
    Tmp tmp; const tops_buf *xs, *ys;
    TRY(contig(ctx, x, tmp, &xs)); TRY(contig(ctx, y, tmp, &ys));
    TRY(prep_out(ctx, out, TOPS_F32, 0, nullptr));
    k::dot(lc_of(ctx), (const float*)xs->data, (const float*)ys->data, xs->numel, (float*)(*out)->data, ctx->scratch);
    return check_launch(ctx, "dot");
}
extern "C" int tops_sum(tops_ctx* ctx, const tops_buf* x, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, x, "tops_sum"));
    TRY(prep_out(ctx, out, TOPS_F32, 0, nullptr));
    k::sum_all(lc_of(ctx), (const float*)x->data, x->numel, (float*)(*out)->data, ctx->scratch);   // order-insensitive: tr irrelevant
    return check_launch(ctx, "sum");
}
extern "C" int tops_ger(tops_ctx* ctx, const tops_buf* x, const tops_buf* y, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, x, "tops_ger")); TRY(need_f32(ctx, y, "tops_ger"));
    if (x->rank != 1 || y->rank != 1) return set_err(ctx, TOPS_ERR_SHAPE, "ger: operands must be vectors");
    int64_t d[2] = {x->dims[0], y->dims[0]};
    TRY(prep_out(ctx, out, TOPS_F32, 2, d));
    k::ger(lc_of(ctx), (const float*)x->data, (const float*)y->data, (float*)(*out)->data, d[0], d[1]);
    return check_launch(ctx, "ger");
}
extern "C" int tops_gemv(tops_ctx* ctx, double alpha, const tops_buf* a, const tops_buf* x, double beta, const tops_buf* y, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, a, "tops_gemv")); TRY(need_f32(ctx, x, "tops_gemv"));
    if (a->rank != 2 || x->rank != 1 || a->dims[1] != x->dims[0]) return set_err(ctx, TOPS_ERR_SHAPE, "gemv: A[%lld,%lld] x[%lld]", (long long)(a->rank == 2 ? a->dims[0] : -1), (long long)(a->rank == 2 ? a->dims[1] : -1), (long long)x->numel);
    if (y) { TRY(need_f32(ctx, y, "tops_gemv")); if (y->numel != a->dims[0]) return set_err(ctx, TOPS_ERR_SHAPE, "gemv: y has %lld elements, expected %lld", (long long)y->numel, (long long)a->dims[0]); }
    int64_t d[1] = {a->dims[0]};
    TRY(prep_out(ctx, out, TOPS_F32, 1, d));
    k::gemv(lc_of(ctx), (float)alpha, (const float*)a->data, a->tr ? 1 : 0, (const float*)x->data, (float)beta, y ? (const float*)y->data : nullptr, (float*)(*out)->data, a->dims[0], a->dims[1]);
    return check_launch(ctx, "gemv");
}
extern "C" int tops_gemm(tops_ctx* ctx, double alpha, const tops_buf* a, const tops_buf* b, double beta, const tops_buf* c, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, a, "tops_gemm")); TRY(need_f32(ctx, b, "tops_gemm"));
    if (a->rank != 2 || b->rank != 2 || a->dims[1] != b->dims[0]) return set_err(ctx, TOPS_ERR_SHAPE, "gemm: inner dimensions do not agree");
    Tmp tmp; const tops_buf* cs = nullptr;
    if (c) {
        TRY(need_f32(ctx, c, "tops_gemm"));
        if (c->rank != 2 || c->dims[0] != a->dims[0] || c->dims[1] != b->dims[1]) return set_err(ctx, TOPS_ERR_SHAPE, "gemm: C has the wrong shape");
        TRY(contig(ctx, c, tmp, &cs));
    }
    int64_t d[2] = {a->dims[0], b->dims[1]};
    TRY(prep_out(ctx, out, TOPS_F32, 2, d));
    GemmCall g{};
    g.dtype = 0; g.M = (int)d[0]; g.N = (int)d[1]; g.K = (int)a->dims[1];
    as_operand(a, 1, &g.A, &g.lda, &g.major_a);
    as_operand(b, 0, &g.B, &g.ldb, &g.major_b);
    g.epi = EPI_STORE; g.alpha = (float)alpha; g.beta = (float)beta; g.tag = "gemm";
    g.out0 = (*out)->data; g.ld_out0 = d[1];
    if (cs) { g.aux0 = cs->data; g.ld_aux0 = d[1]; }
    else split_k_if_skinny(ctx, g);
    return run_gemm(ctx, g);
}
extern "C" int tops_index(tops_ctx* ctx, const tops_buf* x, const int64_t* idx, double* value) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!x || !value || (x->rank > 0 && !idx)) return set_err(ctx, TOPS_ERR_INVALID, "tops_index: NULL argument");
    if (ctx->capturing) return set_err(ctx, TOPS_ERR_INVALID, "tops_index is not recordable (it reads back to the host)");
    int64_t off = 0;
    if (x->tr) {
        if (idx[0] < 0 || idx[0] >= x->dims[0] || idx[1] < 0 || idx[1] >= x->dims[1]) return set_err(ctx, TOPS_ERR_SHAPE, "index out of range");
        off = idx[1] * x->dims[0] + idx[0];
    } else {
        for (int i = 0; i < x->rank; ++i) {
            if (idx[i] < 0 || idx[i] >= x->dims[i]) return set_err(ctx, TOPS_ERR_SHAPE, "index out of range");
            off = off * x->dims[i] + idx[i];
        }
    }
    if (x->dtype == TOPS_F32) {
        float v;
        CUDA_TRY(ctx, cudaMemcpyAsync(&v, (const float*)x->data + off, 4, cudaMemcpyDeviceToHost, ctx->stream));
        TRY(tops_sync(ctx));
        *value = v;
    } else {
        uint16_t h;
        CUDA_TRY(ctx, cudaMemcpyAsync(&h, (const uint16_t*)x->data + off, 2, cudaMemcpyDeviceToHost, ctx->stream));
        TRY(tops_sync(ctx));
        uint32_t u = (uint32_t)h << 16; float v; memcpy(&v, &u, 4);
        *value = v;
    }
    return TOPS_OK;
}
extern "C" int tops_index_row(tops_ctx* ctx, const tops_buf* a, int64_t i, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!a || !out) return set_err(ctx, TOPS_ERR_INVALID, "tops_index_row: NULL argument");
    if (a->rank < 1 || i < 0 || i >= a->dims[0]) return set_err(ctx, TOPS_ERR_SHAPE, "row index out of range");
    Tmp tmp; const tops_buf* s;
    TRY(contig(ctx, a, tmp, &s));
    const int64_t row = s->numel / s->dims[0];
    *out = nullptr;
    return new_buf(ctx, s->dtype, s->rank - 1, s->dims + 1, (char*)s->data + i * row * esize(s->dtype), false, const_cast<tops_buf*>(s), out);
}
extern "C" int tops_transp(tops_ctx* ctx, const tops_buf* a, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!a || !out) return set_err(ctx, TOPS_ERR_INVALID, "tops_transp: NULL argument");
    if (a->rank <= 2 && *out == nullptr) {   // O(1): rank <= 1 is the identity, rank 2 flips the storage flag (hmatrix `tr`, HMat.hs:175)
        int64_t d[2] = {0, 0};
        for (int i = 0; i < a->rank; ++i) d[i] = a->dims[a->rank - 1 - i];
        TRY(new_buf(ctx, a->dtype, a->rank, d, a->data, false, const_cast<tops_buf*>(a), out));
        (*out)->tr = a->rank == 2 ? !a->tr : false;
        return TOPS_OK;
    }
    TRY(need_f32(ctx, a, "tops_transp"));
    int64_t d[TOPS_MAX_RANK]; int perm[TOPS_MAX_RANK];
    for (int i = 0; i < a->rank; ++i) { d[i] = a->dims[a->rank - 1 - i]; perm[i] = a->rank - 1 - i; }
    TRY(prep_out(ctx, out, TOPS_F32, a->rank, d));
    if (a->rank == 2 && a->tr) {   // storage already is the row-major transpose: plain copy
        CUDA_TRY(ctx, cudaMemcpyAsync((*out)->data, a->data, (size_t)a->numel * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        return TOPS_OK;
    }
    if (a->rank == 2) k::transpose2d(lc_of(ctx), (const float*)a->data, (float*)(*out)->data, a->dims[0], a->dims[1]);
    else k::permute(lc_of(ctx), (const float*)a->data, (float*)(*out)->data, a->rank, a->dims, perm);
    return check_launch(ctx, "transp");
}
extern "C" int tops_eye(tops_ctx* ctx, int64_t n, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    int64_t d[2] = {n, n};
    TRY(prep_out(ctx, out, TOPS_F32, 2, d));
    k::eye(lc_of(ctx), (float*)(*out)->data, n);
    return check_launch(ctx, "eye");
}
extern "C" int tops_trace(tops_ctx* ctx, const tops_buf* a, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, a, "tops_trace"));
    if (a->rank != 2 || a->dims[0] != a->dims[1]) return set_err(ctx, TOPS_ERR_SHAPE, "trace: matrix must be square");
    TRY(prep_out(ctx, out, TOPS_F32, 0, nullptr));
    k::trace(lc_of(ctx), (const float*)a->data, a->dims[0], a->dims[0], (float*)(*out)->data);
    return check_launch(ctx, "trace");
}
extern "C" int tops_diag(tops_ctx* ctx, int rank, const tops_buf* v, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, v, "tops_diag"));
    if (v->rank != 1 || rank < 1 || rank > TOPS_MAX_RANK) return set_err(ctx, TOPS_ERR_SHAPE, "diag: need a vector and 1 <= rank <= %d", TOPS_MAX_RANK);
    int64_t d[TOPS_MAX_RANK];
    for (int i = 0; i < rank; ++i) d[i] = v->dims[0];
    TRY(prep_out(ctx, out, TOPS_F32, rank, d));
    k::diag_embed(lc_of(ctx), (const float*)v->data, (float*)(*out)->data, v->dims[0], rank);
    return check_launch(ctx, "diag");
}
extern "C" int tops_get_diag(tops_ctx* ctx, const tops_buf* a, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, a, "tops_get_diag"));
    if (a->rank < 1) return set_err(ctx, TOPS_ERR_SHAPE, "get_diag: rank must be >= 1");
    for (int i = 1; i < a->rank; ++i) if (a->dims[i] != a->dims[0]) return set_err(ctx, TOPS_ERR_SHAPE, "get_diag: all dimensions must agree");
    int64_t d[1] = {a->dims[0]};
    TRY(prep_out(ctx, out, TOPS_F32, 1, d));
    k::diag_extract(lc_of(ctx), (const float*)a->data, (float*)(*out)->data, a->dims[0], a->rank);   // the diagonal is invariant under transposition
    return check_launch(ctx, "get_diag");
}

extern "C" int tops_lift(tops_ctx* ctx, const int32_t* prog, int prog_len, const float* consts, int n_consts,
                         int n_in, const tops_buf* const* in, int rank, const int64_t* dims, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!prog || prog_len <= 0 || prog_len > 64 || n_consts < 0 || n_consts > 16 || n_in < 0 || n_in > 8)
        return set_err(ctx, TOPS_ERR_INVALID, "lift: program too long (max 64 ops, 16 constants, 8 inputs)");
    k::LiftProgram lp{};
    lp.len = prog_len; lp.n_consts = n_consts;
    int depth = 0, maxdepth = 0;
    for (int i = 0; i < prog_len; ++i) {
        const int op = prog[i] >> 16, arg = prog[i] & 0xffff;
        lp.code[i] = prog[i];
        if (op == TOPS_OP_VAR) { if (arg >= n_in) return set_err(ctx, TOPS_ERR_INVALID, "lift: variable %d out of range", arg); ++depth; }
        else if (op == TOPS_OP_CONST) { if (arg >= n_consts) return set_err(ctx, TOPS_ERR_INVALID, "lift: constant %d out of range", arg); ++depth; }
        else if (op == TOPS_OP_ADD || op == TOPS_OP_SUB || op == TOPS_OP_MUL || op == TOPS_OP_DIV || op == TOPS_OP_MAX || op == TOPS_OP_MIN || op == TOPS_OP_POW) { if (depth < 2) return set_err(ctx, TOPS_ERR_INVALID, "lift: stack underflow"); --depth; }
        else if (op >= TOPS_OP_NEG && op <= TOPS_OP_COS) { if (depth < 1) return set_err(ctx, TOPS_ERR_INVALID, "lift: stack underflow"); }
        else return set_err(ctx, TOPS_ERR_INVALID, "lift: unknown opcode %d", op);
        if (depth > maxdepth) maxdepth = depth;
    }
    if (depth != 1 || maxdepth > 16) return set_err(ctx, TOPS_ERR_INVALID, "lift: program must leave exactly one value (stack depth <= 16)");
    for (int i = 0; i < n_consts; ++i) lp.consts[i] = consts[i];
    Tmp tmp;
    const float* ptrs[8] = {};
    int64_t n = 1;
    for (int i = 0; i < rank; ++i) n *= dims[i];
    for (int j = 0; j < n_in; ++j) {
        TRY(need_f32(ctx, in[j], "tops_lift"));
        const tops_buf* s;
        TRY(contig(ctx, in[j], tmp, &s));
        if (s->numel != n) return set_err(ctx, TOPS_ERR_SHAPE, "lift: input %d has %lld elements, expected %lld", j, (long long)s->numel, (long long)n);
        ptrs[j] = (const float*)s->data;
    }
    TRY(prep_out(ctx, out, TOPS_F32, rank, dims));
    k::lift(lc_of(ctx), lp, n_in, ptrs, (float*)(*out)->data, n);
    return check_launch(ctx, "lift");
}

// ================================================================================================ class Tensor
extern "C" int tops_gmul(tops_ctx* ctx, int lM, int lO, int lN, const tops_buf* x, const tops_buf* y, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, x, "tops_gmul")); TRY(need_f32(ctx, y, "tops_gmul"));
    if (lM < 0 || lO < 0 || lN < 0 || x->rank != lM + lO || y->rank != lO + lN || lM + lN > TOPS_MAX_RANK)
        return set_err(ctx, TOPS_ERR_SHAPE, "gmul: ranks %d,%d do not match |ms|=%d |os|=%d |ns|=%d", x->rank, y->rank, lM, lO, lN);
    for (int i = 0; i < lO; ++i)
        if (x->dims[lM + i] != y->dims[lO - 1 - i]) return set_err(ctx, TOPS_ERR_SHAPE, "gmul: contraction dimension %d differs (%lld vs %lld; y's are reversed)", i, (long long)x->dims[lM + i], (long long)y->dims[lO - 1 - i]);
    int64_t od[TOPS_MAX_RANK]; int64_t Mx = 1, O = 1, N = 1;
    for (int i = 0; i < lM; ++i) { od[i] = x->dims[i]; Mx *= x->dims[i]; }
    for (int i = 0; i < lO; ++i) O *= x->dims[lM + i];
    for (int i = 0; i < lN; ++i) { od[lM + i] = y->dims[lO + i]; N *= y->dims[lO + i]; }
    Tmp tmp;
    // Bring y's leading (reversed) contraction axes into x's order: y'[o.., n..].  |os| <= 1 needs nothing.
    const tops_buf* yp = y;
    if (lO >= 2) {
        const tops_buf* yc;
        TRY(contig(ctx, y, tmp, &yc));
        int perm[TOPS_MAX_RANK]; int64_t pd[TOPS_MAX_RANK];
        for (int i = 0; i < lO; ++i) perm[i] = lO - 1 - i;
        for (int i = lO; i < lO + lN; ++i) perm[i] = i;
        for (int i = 0; i < lO + lN; ++i) pd[i] = yc->dims[perm[i]];
        tops_buf* t = nullptr;
        TRY(alloc_buf(ctx, TOPS_F32, lO + lN, pd, &t));
        tmp.keep(t);
        k::permute(lc_of(ctx), (const float*)yc->data, (float*)t->data, lO + lN, yc->dims, perm);
        yp = t;
    }
    // rank-2 transposed views are consumed directly when they are a plain matrix operand; otherwise materialise
    const bool x_mat = (lM == 1 && lO == 1), y_mat = (lO == 1 && lN == 1);
    const tops_buf* xs = x;
    if (x->tr && !x_mat) TRY(contig(ctx, x, tmp, &xs));
    if (yp->tr && !y_mat) TRY(contig(ctx, yp, tmp, &yp));
    TRY(prep_out(ctx, out, TOPS_F32, lM + lN, od));
    float* o = (float*)(*out)->data;
    const float* xd = (const float*)xs->data;
    const float* yd = (const float*)yp->data;
    if (Mx * N == 0) return TOPS_OK;
    if (lO == 0) {                                   // outer product / scaling (dispatchBLAS :146-168, naiveGMul for rank >= 3)
        k::ger(lc_of(ctx), xd, yd, o, Mx, N);
        return check_launch(ctx, "gmul/outer");
    }
    if (Mx == 1 && N == 1) {                         // full contraction -> dot (dispatchDot)
        k::dot(lc_of(ctx), xd, yd, O, o, ctx->scratch);
        return check_launch(ctx, "gmul/dot");
    }
    if (N == 1) {                                    // matrix-vector (dispatchMV)
        if (xs->tr) k::gemv(lc_of(ctx), 1.f, xd, 1, yd, 0.f, nullptr, o, Mx, O);
        else k::gemv(lc_of(ctx), 1.f, xd, 0, yd, 0.f, nullptr, o, Mx, O);
        return check_launch(ctx, "gmul/mv");
    }
    if (Mx == 1) {                                   // vector-matrix (dispatchVM): out[n] = sum_o y'[o,n] x[o]
        if (yp->tr) k::gemv(lc_of(ctx), 1.f, yd, 0, xd, 0.f, nullptr, o, N, O);      // storage is [N,O] row-major
        else k::gemv(lc_of(ctx), 1.f, yd, 1, xd, 0.f, nullptr, o, N, O);
        return check_launch(ctx, "gmul/vm");
    }
    GemmCall g{};                                    // everything else: ONE tensor-core GEMM on flat storage
    g.dtype = 0; g.M = (int)Mx; g.N = (int)N; g.K = (int)O;
    g.A = xd; g.lda = xs->tr ? Mx : O; g.major_a = xs->tr ? MAJOR_MN : MAJOR_K;
    g.B = yd; g.ldb = yp->tr ? O : N; g.major_b = yp->tr ? MAJOR_K : MAJOR_MN;
    g.epi = EPI_STORE; g.alpha = 1.f; g.beta = 0.f; g.out0 = o; g.ld_out0 = N; g.tag = "gmul";
    split_k_if_skinny(ctx, g);
    return run_gemm(ctx, g);
}

// `gmul lM lO lN >>> sumRows` as one primitive (a fusion the deferred evaluator applies to the composed TOp, top.py): the row sum
// commutes with the contraction, so  sumRows (gmul x y) = gmul (sumRows x) y  and the [A, ...] intermediate is never formed.
//   lO == 1, small y: ONE pass over x (k_gsr_fwd).   Otherwise: sum_rows + gmul on the reduced operand.
namespace {
bool gsr_shapes(tops_ctx* ctx, int lM, int lO, int lN, const tops_buf* x, const tops_buf* y, int64_t* A, int64_t* R, int64_t* K, int64_t* N) {
    if (lO != 1 || x->tr || y->tr || x->numel == 0 || y->numel == 0) return false;
    *A = x->dims[0]; *R = 1; *N = 1;
    for (int i = 1; i < lM; ++i) *R *= x->dims[i];
    *K = x->dims[lM];
    for (int i = 0; i < lN; ++i) *N *= y->dims[lO + i];
    return k::gsr_fits(*K, *N) && *R < (1ll << 31);
}
int gsr_check(tops_ctx* ctx, int lM, int lO, int lN, const tops_buf* x, const tops_buf* y) {
    TRY(need_f32(ctx, x, "tops_gmul_sum_rows")); TRY(need_f32(ctx, y, "tops_gmul_sum_rows"));
    if (lM < 1 || lO < 0 || lN < 0 || x->rank != lM + lO || y->rank != lO + lN || lM - 1 + lN > TOPS_MAX_RANK)
        return set_err(ctx, TOPS_ERR_SHAPE, "gmul_sum_rows: ranks %d,%d do not match |ms|=%d (>= 1) |os|=%d |ns|=%d", x->rank, y->rank, lM, lO, lN);
    for (int i = 0; i < lO; ++i)
        if (x->dims[lM + i] != y->dims[lO - 1 - i]) return set_err(ctx, TOPS_ERR_SHAPE, "gmul_sum_rows: contraction dimension %d differs", i);
    return TOPS_OK;
}
}  // namespace

extern "C" int tops_gmul_sum_rows(tops_ctx* ctx, int lM, int lO, int lN, const tops_buf* x, const tops_buf* y, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(gsr_check(ctx, lM, lO, lN, x, y));
    int64_t A, R, K, N;
    if (gsr_shapes(ctx, lM, lO, lN, x, y, &A, &R, &K, &N)) {
        int64_t od[TOPS_MAX_RANK];
        for (int i = 1; i < lM; ++i) od[i - 1] = x->dims[i];
        for (int i = 0; i < lN; ++i) od[lM - 1 + i] = y->dims[lO + i];
        TRY(prep_out(ctx, out, TOPS_F32, lM - 1 + lN, od));
        ProfScope prof_(ctx, "gmul_sum_rows", 2.0 * R * K * N, 4.0 * ((double)x->numel + y->numel + R * N));
        k::gsr_fwd(lc_of(ctx), (const float*)x->data, (const float*)y->data, (float*)(*out)->data, A, R, (int)K, (int)N);
        return check_launch(ctx, "gmul_sum_rows");
    }
    tops_buf* xs = nullptr;
    TRY(tops_sum_rows(ctx, x, &xs));
    Tmp tmp; tmp.keep(xs);
    return tops_gmul(ctx, lM - 1, lO, lN, xs, y, out);
}

extern "C" int tops_gmul_sum_rows_vjp(tops_ctx* ctx, int lM, int lO, int lN, const tops_buf* x, const tops_buf* y, const tops_buf* ct,
                                      tops_buf** dx, tops_buf** dy) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(gsr_check(ctx, lM, lO, lN, x, y));
    TRY(need_f32(ctx, ct, "tops_gmul_sum_rows_vjp"));
    if (!dx || !dy) return set_err(ctx, TOPS_ERR_INVALID, "gmul_sum_rows_vjp: NULL output slot");
    int64_t A, R, K, N;
    if (gsr_shapes(ctx, lM, lO, lN, x, y, &A, &R, &K, &N) && !ct->tr) {
        if (ct->numel != R * N) return set_err(ctx, TOPS_ERR_SHAPE, "gmul_sum_rows_vjp: cotangent has %lld elements, expected %lld", (long long)ct->numel, (long long)(R * N));
        TRY(prep_out(ctx, dx, TOPS_F32, x->rank, x->dims));
        TRY(prep_out(ctx, dy, TOPS_F32, y->rank, y->dims));
        ProfScope prof_(ctx, "gmul_sum_rows_vjp", 4.0 * R * K * N, 4.0 * (2.0 * x->numel + 2.0 * y->numel + R * N));
        if (k::gsr_vjp_needs_zeroed_dy(R, (int)K, (int)N)) CUDA_TRY(ctx, cudaMemsetAsync((*dy)->data, 0, sizeof(float) * (size_t)(K * N), ctx->stream));
        k::gsr_vjp(lc_of(ctx), (const float*)x->data, (const float*)y->data, (const float*)ct->data, (float*)(*dx)->data, (float*)(*dy)->data, A, R, (int)K, (int)N);
        return check_launch(ctx, "gmul_sum_rows_vjp");
    }
    // general shapes: the reference's own VJP chain (TOp.hs:73-93,151-159) on the existing primitives
    Tmp tmp;
    tops_buf *dz = nullptr, *yt = nullptr, *xt = nullptr;
    TRY(tops_broadcast_rows(ctx, x->dims[0], ct, &dz)); tmp.keep(dz);
    TRY(tops_transp(ctx, y, &yt)); tmp.keep(yt);
    TRY(tops_transp(ctx, x, &xt)); tmp.keep(xt);
    TRY(tops_gmul(ctx, lM, lN, lO, dz, yt, dx));
    return tops_gmul(ctx, lO, lM, lN, xt, dz, dy);
}

extern "C" int tops_sum_t(tops_ctx* ctx, int n, const tops_buf* const* xs, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (n < 1 || !xs) return set_err(ctx, TOPS_ERR_INVALID, "sum_t: need at least one tensor");
    Tmp tmp;
    std::vector<const float*> ptrs;
    const tops_buf* first = nullptr;
    for (int j = 0; j < n; ++j) {
        TRY(need_f32(ctx, xs[j], "tops_sum_t"));
        const tops_buf* s;
        TRY(contig(ctx, xs[j], tmp, &s));
        if (!first) first = s;
        else if (!same_shape(first, s)) return set_err(ctx, TOPS_ERR_SHAPE, "sum_t: operand %d has a different shape", j);
        ptrs.push_back((const float*)s->data);
    }
    TRY(prep_out(ctx, out, TOPS_F32, first->rank, first->dims));
    float* o = (float*)(*out)->data;
    // left fold in order, 8 operands per pass
    int done = 0;
    while (done < n) {
        const float* batch[8]; int nb = 0;
        if (done > 0) batch[nb++] = o;
        while (nb < 8 && done < n) batch[nb++] = ptrs[done++];
        k::add_n(lc_of(ctx), nb, batch, o, first->numel);
    }
    return check_launch(ctx, "sum_t");
}
extern "C" int tops_sum_rows(tops_ctx* ctx, const tops_buf* x, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, x, "tops_sum_rows"));
    if (x->rank < 1) return set_err(ctx, TOPS_ERR_SHAPE, "sum_rows: rank must be >= 1");
    Tmp tmp; const tops_buf* s;
    TRY(contig(ctx, x, tmp, &s));
    const int64_t rows = s->dims[0], cols = rows ? s->numel / rows : 0;
    TRY(prep_out(ctx, out, TOPS_F32, s->rank - 1, s->dims + 1));
    if (rows == 0) { k::fill(lc_of(ctx), (float*)(*out)->data, (*out)->numel, 0.f); return check_launch(ctx, "sum_rows"); }
    float* ws = nullptr;
    TRY(ws_alloc(ctx, sizeof(float) * 64 * (size_t)(cols > 0 ? cols : 1), (void**)&ws));
    k::col_sums(lc_of(ctx), (const float*)s->data, rows, cols, (float*)(*out)->data, ws);
    ws_free(ctx, ws);
    return check_launch(ctx, "sum_rows");
}
extern "C" int tops_broadcast_rows(tops_ctx* ctx, int64_t n, const tops_buf* row, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, row, "tops_broadcast_rows"));
    if (row->rank + 1 > TOPS_MAX_RANK || n < 0) return set_err(ctx, TOPS_ERR_SHAPE, "broadcast_rows: bad shape");
    Tmp tmp; const tops_buf* s;
    TRY(contig(ctx, row, tmp, &s));
    int64_t d[TOPS_MAX_RANK]; d[0] = n;
    for (int i = 0; i < s->rank; ++i) d[i + 1] = s->dims[i];
    TRY(prep_out(ctx, out, TOPS_F32, s->rank + 1, d));
    k::broadcast_rows(lc_of(ctx), (const float*)s->data, (float*)(*out)->data, n, s->numel);
    return check_launch(ctx, "broadcast_rows");
}
extern "C" int tops_map_rows_softmax(tops_ctx* ctx, const tops_buf* x, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(need_f32(ctx, x, "tops_map_rows_softmax"));
    if (x->rank != 2) return set_err(ctx, TOPS_ERR_SHAPE, "map_rows_softmax: need a [rows, cols] matrix");
    Tmp tmp; const tops_buf* s;
    TRY(contig(ctx, x, tmp, &s));
    TRY(prep_out(ctx, out, TOPS_F32, 2, s->dims));
    k::softmax_rows(lc_of(ctx), (const float*)s->data, (float*)(*out)->data, s->dims[0], s->dims[1]);
    return check_launch(ctx, "softmax");
}

// ================================================================================================ fused batched hot path
namespace {

struct LayerShapes { int64_t B, i, o; int dtype; };

int layer_shapes(tops_ctx* ctx, const tops_buf* X, const tops_buf* W, const tops_buf* b, LayerShapes* s) {
    if (!X || !W) return set_err(ctx, TOPS_ERR_INVALID, "fflayer: NULL tensor");
    if (X->tr || W->tr) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "fflayer: transposed views are not accepted");
    if (X->rank != 2 || W->rank != 2 || X->dims[1] != W->dims[1]) return set_err(ctx, TOPS_ERR_SHAPE, "fflayer: X[B,i] W[o,i] expected (FeedForward.hs:209-212)");
    if (b && (b->rank != 1 || b->dims[0] != W->dims[0] || b->dtype != TOPS_F32)) return set_err(ctx, TOPS_ERR_SHAPE, "fflayer: b[o] (fp32) expected");
    if (X->dtype != W->dtype) return set_err(ctx, TOPS_ERR_SHAPE, "fflayer: X and W must have the same dtype");
    s->B = X->dims[0]; s->i = X->dims[1]; s->o = W->dims[0]; s->dtype = X->dtype;
    return TOPS_OK;
}

// TOPS_PREC_TF32_BF16X2: the small operand W takes part in two GEMMs of a layer (forward and dX); its bf16 correction operands
// bf16(W), bf16(W - trunc_tf32(W)) are produced once per call in HBM (4 extra bytes per weight) so that the GEMMs' splitter warps
// only re-tile the big operand.  Leaves NULL pointers when the mode / shape does not apply.
struct WSplit { const void* w16 = nullptr; const void* wlo16 = nullptr; };
int make_wsplit(tops_ctx* ctx, const tops_buf* W, Tmp& tmp, WSplit* out) {
    if (ctx->precision != TOPS_PREC_TF32_BF16X2 || W->dtype != TOPS_F32 || W->tr || W->rank != 2 || (W->dims[1] % 8) != 0 || W->numel == 0) return TOPS_OK;
    tops_buf *a = nullptr, *b = nullptr;
    TRY(alloc_buf(ctx, TOPS_BF16, 2, W->dims, &a)); tmp.keep(a);
    TRY(alloc_buf(ctx, TOPS_BF16, 2, W->dims, &b)); tmp.keep(b);
    k::split_bf16(lc_of(ctx), (const float*)W->data, a->data, b->data, W->numel);
    TRY(check_launch(ctx, "split_bf16"));
    out->w16 = a->data; out->wlo16 = b->data;
    return TOPS_OK;
}

// A = act(X W^T + b)  [optionally also dZ = dA ⊙ act'(A)]  — one GEMM, everything else in its epilogue
// `db` (optional, epilogues with a dZ output only): column sums of dZ fused into the epilogue; *db_fused reports whether the
// kernel produced them (the TMA epilogue does; the direct / SIMT paths leave it to col_sums).
int fwd_gemm(tops_ctx* ctx, const LayerShapes& s, const void* X, const void* W, const float* b, int act, int epi,
             void* A, const void* aux, void* out1, float* loss, float* db = nullptr, int* db_fused = nullptr, bool db_accumulate = false,
             const WSplit* ws = nullptr) {
    GemmCall g{};
    g.dtype = s.dtype == TOPS_BF16; g.M = (int)s.B; g.N = (int)s.o; g.K = (int)s.i;
    g.A = X; g.lda = s.i; g.major_a = MAJOR_K;
    g.B = W; g.ldb = s.i; g.major_b = MAJOR_K;
    if (ws) { g.B16 = ws->w16; g.Blo16 = ws->wlo16; }
    g.epi = epi; g.act = act; g.alpha = 1.f; g.bias = b; g.tag = "gemm_fwd";
    g.out0 = A; g.ld_out0 = s.o; g.out1 = out1; g.ld_out1 = s.o; g.aux0 = aux; g.ld_aux0 = s.o; g.loss = loss;
    g.io_bf16 = g.dtype;
    // 3xTF32: a heavy fused epilogue (aux operand + two outputs) keeps the epilogue warps away from draining TMEM chunks; chunks
    // of 8 k-blocks let the MMA warp run 16 k-blocks ahead meanwhile (GEMM error 1.8e-6 / 2.4e-6 instead of 1.35e-6 / 1.5e-6 in the
    // TF32_BF16X2 / TF32X3 modes, bar 1e-5; chunks of 4 cost 12 % of this GEMM's time for 10 % less gradient error: measured)
    if (aux != nullptr) g.chunk_kb = ctx->fused_chunk_kb;
    if (db && out1 && s.B > 0) {
        if (!db_accumulate) CUDA_TRY(ctx, cudaMemsetAsync(db, 0, sizeof(float) * (size_t)s.o, ctx->stream));
        g.colsum = db; g.colsum_src = 2; g.colsum_fused = db_fused;
    }
    return run_gemm(ctx, g);
}
// dX = dZ W  (optionally ⊙ act'(A_prev))
// `db_prev` (optional, EPI_MUL_DACT only): the output IS dZ of the previous layer, so its column sums are that layer's db.
int dx_gemm(tops_ctx* ctx, const LayerShapes& s, const void* dZ, const void* W, void* dX, int epi, int act, const void* Aprev,
            float* db_prev = nullptr, int* db_fused = nullptr, const WSplit* ws = nullptr, int max_ctas = 0) {
    GemmCall g{};
    g.dtype = s.dtype == TOPS_BF16; g.M = (int)s.B; g.N = (int)s.i; g.K = (int)s.o;
    g.A = dZ; g.lda = s.o; g.major_a = MAJOR_K;
    g.B = W; g.ldb = s.i; g.major_b = MAJOR_MN;
    if (ws) { g.B16 = ws->w16; g.Blo16 = ws->wlo16; }
    g.epi = epi; g.act = act; g.alpha = 1.f; g.tag = "gemm_dX";
    g.out0 = dX; g.ld_out0 = s.i; g.aux0 = Aprev; g.ld_aux0 = s.i;
    g.io_bf16 = g.dtype; g.max_ctas = max_ctas;
    if (db_prev && epi == EPI_MUL_DACT && s.B > 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(db_prev, 0, sizeof(float) * (size_t)s.i, ctx->stream));
        g.colsum = db_prev; g.colsum_src = 1; g.colsum_fused = db_fused;
    }
    return run_gemm(ctx, g);
}
// dW = dZ^T Xin   (split-K over the batch, fp32 atomics into a zeroed output), db = column sums of dZ
int dw_db(tops_ctx* ctx, const LayerShapes& s, const void* dZ, const void* Xin, float* dW, float* db, bool db_done = false, bool accumulate = false,
          float* dW_mc = nullptr) {
    GemmCall g{};
    g.dtype = s.dtype == TOPS_BF16; g.M = (int)s.o; g.N = (int)s.i; g.K = (int)s.B;
    g.A = dZ; g.lda = s.o; g.major_a = MAJOR_MN;
    g.B = Xin; g.ldb = s.i; g.major_b = MAJOR_MN;
    g.epi = EPI_ATOMIC; g.alpha = 1.f; g.out0 = dW; g.ld_out0 = s.i; g.tag = "gemm_dW"; g.accumulate = accumulate ? 1 : 0;
    int* counters = nullptr;
    if (dW_mc) {   // fused all-reduce: region-arrival counters for the "last split pushes the finished region" protocol
        const size_t n = ((size_t)(s.o + 127) / 128 + 1) * ((size_t)(s.i + 127) / 128 + 1) * 8 + 16;   // >= tiles * CTAs per tile * 8 epilogue warps
        TRY(ws_alloc(ctx, n * sizeof(int), (void**)&counters));
        CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, n * sizeof(int), ctx->stream));
        g.out0_mc = dW_mc; g.tile_counters = counters;
    }
    int rc_ = run_gemm(ctx, g);
    ws_free(ctx, counters);
    TRY(rc_);
    if (db && !db_done) {   // the forward epilogue could not fuse the column sums (rows not 16-byte aligned, FP32_SIMT, SIMT fallback)
        ProfScope prof_(ctx, "col_sums_db", 0.0, (s.dtype == TOPS_BF16 ? 2.0 : 4.0) * (double)s.B * s.o);
        float* ws = nullptr;
        TRY(ws_alloc(ctx, sizeof(float) * 64 * (size_t)s.o, (void**)&ws));
        if (s.dtype == TOPS_BF16) k::col_sums_bf16(lc_of(ctx), dZ, s.B, s.o, db, ws, accumulate);
        else k::col_sums(lc_of(ctx), (const float*)dZ, s.B, s.o, db, ws, accumulate);
        ws_free(ctx, ws);
        TRY(check_launch(ctx, "col_sums"));
    }
    return TOPS_OK;
}

// ---- TOPS_PREC_F16X3: the whole forward + VJP of one layer on fp16 pairs, no operand split more than once.
//   W  -> (W1, W2)  one scale sW for the tensor            (absmax + split: 2 small launches per call)
//   X  -> (X1, X2)  one scale per sample row, rsX[s] = 1/scale; the same launch reduces max |dA|   (the only extra passes over HBM)
//   forward GEMM    Z = acc * rsX[s] / sW + b, A = act(Z), dZ = dA * act'(A); its epilogue emits A (fp32), db (column sums of the
//                   fp32 dZ) and dZ directly as the pair (dZ1, dZ2) of dZ * c * rsX[s] — fp32 dZ never exists in HBM
//   dW GEMM         sum_s dZ'[s,o] X'[s,i] = c * dW: the per-sample factors cancel inside the contraction
//   dX GEMM         dX[s,:] = acc / (c sW rsX[s])
// c = 2^(13 - e(max|dA|)) / max_s rsX[s] keeps |dZ'| < 2^14 (|act'| <= 1).  All factors are powers of two: the scaling is exact.
// fp16 pair of a layer's W.  `pending`: the planes are allocated but not written yet — the split rides along with the two launches
// of the X row split inside layer_fwd_grad_f16x3 (max|W| with the rows, the pair with the fix-up pass) instead of two of its own.
struct WPairF16 { void* w1 = nullptr; void* w2 = nullptr; float* s2 = nullptr; bool pending = false; const float* src = nullptr; int64_t n = 0; unsigned* mx = nullptr; };

bool f16x3_layer_ok(tops_ctx* ctx, const LayerShapes& s) {
    return ctx->precision == TOPS_PREC_F16X3 && s.dtype == TOPS_F32 && s.B > 0 && s.i > 0 && s.o > 0 && (s.i % 8) == 0 && (s.o % 8) == 0;
}

int prep_wpair_f16(tops_ctx* ctx, const LayerShapes& s, const void* W, Tmp& tmp, WPairF16* out, bool defer = false) {
    int64_t pd[1] = {s.o * s.i}, sd[1] = {4};
    tops_buf *hi = nullptr, *lo = nullptr, *sc = nullptr;
    TRY(alloc_buf(ctx, TOPS_BF16, 1, pd, &hi)); tmp.keep(hi);
    TRY(alloc_buf(ctx, TOPS_BF16, 1, pd, &lo)); tmp.keep(lo);
    TRY(alloc_buf(ctx, TOPS_F32, 1, sd, &sc)); tmp.keep(sc);
    unsigned* mx = reinterpret_cast<unsigned*>(sc->data) + 2;
    CUDA_TRY(ctx, cudaMemsetAsync(mx, 0, 4, ctx->stream));
    if (defer) {
        *out = WPairF16{hi->data, lo->data, (float*)sc->data, true, (const float*)W, s.o * s.i, mx};
        return TOPS_OK;
    }
    ProfScope prof_(ctx, "split_f16_W", 0.0, 12.0 * (double)s.o * s.i);
    k::absmax_bits(lc_of(ctx), (const float*)W, s.o * s.i, mx);
    k::split_f16_tensor(lc_of(ctx), (const float*)W, s.o * s.i, mx, hi->data, lo->data, (float*)sc->data);
    TRY(check_launch(ctx, "split_f16(W)"));
    out->w1 = hi->data; out->w2 = lo->data; out->s2 = (float*)sc->data;
    return TOPS_OK;
}

// X, dA, A, dX: device pointers to [B,i] / [B,o] fp32 row-major; dW [o,i], db [o] (nullable) fp32.  accumulate: dW/db are added to.
// ev_grads (nullable): recorded on the stream as soon as dW and db are complete, i.e. BEFORE the dX GEMM is launched — a
// data-parallel caller starts the all-reduce of [dW‖db] on another stream while dX runs on at most dx_max_ctas SMs (0 = all).
int layer_fwd_grad_f16x3(tops_ctx* ctx, const LayerShapes& s, const void* X, const WPairF16& wp, const float* b, int act, const void* dA,
                         void* A, void* dX, float* dW, float* db, bool accumulate, float* dW_mc, cudaEvent_t ev_grads = nullptr, int dx_max_ctas = 0) {
    Tmp tmp;
    int64_t xd[1] = {s.B * s.i}, zd[1] = {s.B * s.o}, rd[1] = {s.B}, sd[1] = {8};
    tops_buf *x1 = nullptr, *x2 = nullptr, *z1 = nullptr, *z2 = nullptr, *rs = nullptr, *sc = nullptr;
    TRY(alloc_buf(ctx, TOPS_BF16, 1, xd, &x1)); tmp.keep(x1);
    TRY(alloc_buf(ctx, TOPS_BF16, 1, xd, &x2)); tmp.keep(x2);
    TRY(alloc_buf(ctx, TOPS_BF16, 1, zd, &z1)); tmp.keep(z1);
    TRY(alloc_buf(ctx, TOPS_BF16, 1, zd, &z2)); tmp.keep(z2);
    TRY(alloc_buf(ctx, TOPS_F32, 1, rd, &rs)); tmp.keep(rs);
    TRY(alloc_buf(ctx, TOPS_F32, 1, sd, &sc)); tmp.keep(sc);   // [0..3] {1/sW, c, 1/c, 1/(c sW)}   [4] max rsX bits   [5] max|dA| bits
    float* scal = (float*)sc->data;
    unsigned* words = reinterpret_cast<unsigned*>(scal) + 4;
    const float* rsX = (const float*)rs->data;
    CUDA_TRY(ctx, cudaMemsetAsync(words, 0, 8, ctx->stream));
    {
        ProfScope prof_(ctx, "split_f16_X", 0.0, 8.0 * (double)s.B * s.i + 4.0 * (double)s.B * s.o + (wp.pending ? 12.0 * (double)s.o * s.i : 0.0));
        // two launches: rows of X (+ max|dA|, + max|W|), then the fix-up pass (+ the pair of W, + the layer's scalars)
        k::SideTensor side{wp.pending ? wp.src : nullptr, wp.n, wp.mx, wp.w1, wp.w2, wp.s2};
        k::split_f16_rows(lc_of(ctx), (const float*)X, s.B, s.i, x1->data, x2->data, (float*)rs->data, words, (const float*)dA, s.o, words + 1,
                          wp.pending ? &side : nullptr, wp.s2, scal);
        TRY(check_launch(ctx, "split_f16(X)"));
    }
    // ---- forward: A, db, (dZ1, dZ2)
    int db_fused = 0;
    {
        GemmCall g{};
        g.dtype = 2; g.M = (int)s.B; g.N = (int)s.o; g.K = (int)s.i;
        g.A = x1->data; g.A2 = x2->data; g.lda = s.i; g.major_a = MAJOR_K;
        g.B = wp.w1; g.B2 = wp.w2; g.ldb = s.i; g.major_b = MAJOR_K;
        g.epi = EPI_BIAS_ACT_DZ; g.act = act; g.alpha = 1.f; g.bias = b; g.tag = "gemm_fwd";
        g.out0 = A; g.ld_out0 = s.o; g.aux0 = dA; g.ld_aux0 = s.o;
        g.out1 = z1->data; g.out1b = z2->data; g.ld_out1 = s.o; g.out1_pair = 1; g.out1_scale_ptr = scal + 1; g.out1_row_scale = rsX;
        g.acc_scale_ptr = scal + 0; g.row_scale = rsX; g.row_scale_inv = 0;
        g.chunk_kb = ctx->f16x3_fwd_chunk_kb; g.chunk_head_kb = ctx->f16x3_fwd_head_kb;
        if (db) {
            if (!accumulate) CUDA_TRY(ctx, cudaMemsetAsync(db, 0, sizeof(float) * (size_t)s.o, ctx->stream));
            g.colsum = db; g.colsum_src = 2; g.colsum_fused = &db_fused;
        }
        TRY(run_gemm(ctx, g));
        if (db && !db_fused) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "F16X3 layer: the staged epilogue was not available for db");
    }
    // ---- dW = dZ^T X (split-K over the batch)
    {
        GemmCall g{};
        g.dtype = 2; g.M = (int)s.o; g.N = (int)s.i; g.K = (int)s.B;
        g.A = z1->data; g.A2 = z2->data; g.lda = s.o; g.major_a = MAJOR_MN;
        g.B = x1->data; g.B2 = x2->data; g.ldb = s.i; g.major_b = MAJOR_MN;
        g.epi = EPI_ATOMIC; g.alpha = 1.f; g.out0 = dW; g.ld_out0 = s.i; g.tag = "gemm_dW"; g.accumulate = accumulate ? 1 : 0;
        g.acc_scale_ptr = scal + 2; g.chunk_kb = ctx->f16x3_chunk_kb;
        int* counters = nullptr;
        if (dW_mc) {
            const size_t n = ((size_t)(s.o + 127) / 128 + 1) * ((size_t)(s.i + 127) / 128 + 1) * 8 + 16;
            TRY(ws_alloc(ctx, n * sizeof(int), (void**)&counters));
            CUDA_TRY(ctx, cudaMemsetAsync(counters, 0, n * sizeof(int), ctx->stream));
            g.out0_mc = dW_mc; g.tile_counters = counters;
        }
        const int rc_ = run_gemm(ctx, g);
        ws_free(ctx, counters);
        TRY(rc_);
    }
    if (ev_grads) CUDA_TRY(ctx, cudaEventRecord(ev_grads, ctx->stream));
    // ---- dX = dZ W
    if (dX) {
        GemmCall g{};
        g.dtype = 2; g.M = (int)s.B; g.N = (int)s.i; g.K = (int)s.o;
        g.A = z1->data; g.A2 = z2->data; g.lda = s.o; g.major_a = MAJOR_K;
        g.B = wp.w1; g.B2 = wp.w2; g.ldb = s.i; g.major_b = MAJOR_MN;
        g.epi = EPI_STORE; g.alpha = 1.f; g.out0 = dX; g.ld_out0 = s.i; g.tag = "gemm_dX";
        g.acc_scale_ptr = scal + 3; g.row_scale = rsX; g.row_scale_inv = 1; g.chunk_kb = ctx->f16x3_chunk_kb;
        g.max_ctas = dx_max_ctas;
        TRY(run_gemm(ctx, g));
    }
    return TOPS_OK;
}

int check_act(tops_ctx* ctx, int act) {
    if (act != TOPS_ACT_ID && act != TOPS_ACT_LOGISTIC) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "fflayer: activation %d is not fusable here (use tops_mlp_* for softmax)", act);
    return TOPS_OK;
}

}  // namespace

extern "C" int tops_fflayer_fwd(tops_ctx* ctx, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act, tops_buf** A) {
    CHECK_CTX(ctx); LOCK(ctx);
    LayerShapes s; TRY(layer_shapes(ctx, X, W, b, &s)); TRY(check_act(ctx, act));
    int64_t d[2] = {s.B, s.o};
    TRY(prep_out(ctx, A, s.dtype, 2, d));
    Tmp tmp; WSplit ws; TRY(make_wsplit(ctx, W, tmp, &ws));
    return fwd_gemm(ctx, s, X->data, W->data, b ? (const float*)b->data : nullptr, act, EPI_BIAS_ACT, (*A)->data, nullptr, nullptr, nullptr, nullptr, nullptr, false, &ws);
}

extern "C" int tops_fflayer_fwd_grad(tops_ctx* ctx, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act,
                                     const tops_buf* dA, tops_buf** A, tops_buf** dX, tops_buf** dW, tops_buf** db) {
    CHECK_CTX(ctx); LOCK(ctx);
    LayerShapes s; TRY(layer_shapes(ctx, X, W, b, &s)); TRY(check_act(ctx, act));
    if (!dA || dA->rank != 2 || dA->dims[0] != s.B || dA->dims[1] != s.o || dA->dtype != s.dtype || dA->tr) return set_err(ctx, TOPS_ERR_SHAPE, "fflayer: dA[B,o] expected");
    int64_t dAo[2] = {s.B, s.o}, dXs[2] = {s.B, s.i}, dWs[2] = {s.o, s.i}, dbs[1] = {s.o};
    TRY(prep_out(ctx, A, s.dtype, 2, dAo));
    if (dX) TRY(prep_out(ctx, dX, s.dtype, 2, dXs));
    TRY(prep_out(ctx, dW, TOPS_F32, 2, dWs));
    if (db) TRY(prep_out(ctx, db, TOPS_F32, 1, dbs));
    Tmp tmp;
    if (f16x3_layer_ok(ctx, s)) {
        WPairF16 wp; TRY(prep_wpair_f16(ctx, s, W->data, tmp, &wp, true));
        return layer_fwd_grad_f16x3(ctx, s, X->data, wp, b ? (const float*)b->data : nullptr, act, dA->data, (*A)->data, dX ? (*dX)->data : nullptr,
                                    (float*)(*dW)->data, db ? (float*)(*db)->data : nullptr, false, nullptr);
    }
    SplitScope split_scope_(ctx);
    tops_buf* dZ = nullptr;
    TRY(alloc_buf(ctx, s.dtype, 2, dAo, &dZ)); tmp.keep(dZ);
    WSplit ws; TRY(make_wsplit(ctx, W, tmp, &ws));
    int db_fused = 0;
    TRY(fwd_gemm(ctx, s, X->data, W->data, b ? (const float*)b->data : nullptr, act, EPI_BIAS_ACT_DZ, (*A)->data, dA->data, dZ->data, nullptr,
                 db ? (float*)(*db)->data : nullptr, &db_fused, false, &ws));
    TRY(dw_db(ctx, s, dZ->data, X->data, (float*)(*dW)->data, db ? (float*)(*db)->data : nullptr, db_fused != 0));
    if (dX) TRY(dx_gemm(ctx, s, dZ->data, W->data, (*dX)->data, EPI_STORE, ACT_ID, nullptr, nullptr, nullptr, &ws));
    return TOPS_OK;
}

// Host-buffer entry point (what the reference's `fromList`-batch -> netGrad -> `toList` round trip becomes): X and dA live in
// HOST memory.  The batch is cut into row chunks; chunk k+1 crosses PCIe on the copy stream while chunk k runs its three
// GEMMs on the compute stream, dW/db accumulating across chunks (split-K atomics / fused column sums).  PCIe is the bound.
extern "C" int tops_fflayer_fwd_grad_host(tops_ctx* ctx, const float* X_host, const float* dA_host, int64_t B, const tops_buf* W, const tops_buf* b,
                                          int act, int n_chunks, tops_buf** A, tops_buf** dX, tops_buf** grads, float* grads_host) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!X_host || !dA_host || !W || !grads || B < 0) return set_err(ctx, TOPS_ERR_INVALID, "fflayer_fwd_grad_host: NULL argument");
    if (ctx->capturing) return set_err(ctx, TOPS_ERR_INVALID, "the host-buffer entry point is not recordable (it uses its own copy stream)");
    if (W->rank != 2 || W->dtype != TOPS_F32 || W->tr || (b && (b->rank != 1 || b->dims[0] != W->dims[0] || b->dtype != TOPS_F32)))
        return set_err(ctx, TOPS_ERR_SHAPE, "fflayer_fwd_grad_host: W[o,i] b[o] fp32 expected");
    TRY(check_act(ctx, act));
    const int64_t o = W->dims[0], i = W->dims[1];
    int64_t dA_[2] = {B, o}, dX_[2] = {B, i}, g_[1] = {o * i + o};
    Tmp tmp;
    SplitScope split_scope_(ctx);
    tops_buf *Xd = nullptr, *dAd = nullptr, *dZ = nullptr, *Ad = nullptr, *dXd = nullptr;
    TRY(alloc_buf(ctx, TOPS_F32, 2, dX_, &Xd)); tmp.keep(Xd);
    TRY(alloc_buf(ctx, TOPS_F32, 2, dA_, &dAd)); tmp.keep(dAd);
    TRY(alloc_buf(ctx, TOPS_F32, 2, dA_, &dZ)); tmp.keep(dZ);
    if (A) { TRY(prep_out(ctx, A, TOPS_F32, 2, dA_)); Ad = *A; } else { TRY(alloc_buf(ctx, TOPS_F32, 2, dA_, &Ad)); tmp.keep(Ad); }
    if (dX) { TRY(prep_out(ctx, dX, TOPS_F32, 2, dX_)); dXd = *dX; } else { TRY(alloc_buf(ctx, TOPS_F32, 2, dX_, &dXd)); tmp.keep(dXd); }
    TRY(prep_out(ctx, grads, TOPS_F32, 1, g_));
    float* dW = (float*)(*grads)->data; float* db = dW + o * i;
    CUDA_TRY(ctx, cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)(o * i + o), ctx->stream));
    if (n_chunks <= 0) n_chunks = 8;
    if (n_chunks > 64) n_chunks = 64;
    int64_t rows_per = ((B + n_chunks - 1) / n_chunks + 127) / 128 * 128;   // whole 128-row tiles per chunk
    if (rows_per <= 0) rows_per = 128;
    // the allocations above are stream-ordered on the compute stream: the copy stream must not write before they exist
    cudaEvent_t ready;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    cudaEventRecord(ready, ctx->stream);
    cudaStreamWaitEvent(ctx->copy_stream, ready, 0);
    std::vector<cudaEvent_t> ev;
    for (int64_t r0 = 0; r0 < B; r0 += rows_per) {
        const int64_t n = (B - r0 < rows_per) ? B - r0 : rows_per;
        cudaMemcpyAsync((float*)Xd->data + r0 * i, X_host + r0 * i, sizeof(float) * (size_t)(n * i), cudaMemcpyHostToDevice, ctx->copy_stream);
        cudaMemcpyAsync((float*)dAd->data + r0 * o, dA_host + r0 * o, sizeof(float) * (size_t)(n * o), cudaMemcpyHostToDevice, ctx->copy_stream);
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        cudaEventRecord(e, ctx->copy_stream);
        ev.push_back(e);
    }
    WSplit ws;
    int rc = make_wsplit(ctx, W, tmp, &ws);
    const LayerShapes s_all{B, i, o, TOPS_F32};
    const bool f16x3 = f16x3_layer_ok(ctx, s_all);
    WPairF16 wp;
    if (rc == TOPS_OK && f16x3) rc = prep_wpair_f16(ctx, s_all, W->data, tmp, &wp);
    int64_t r0 = 0;
    for (size_t c = 0; c < ev.size() && rc == TOPS_OK; ++c, r0 += rows_per) {
        const int64_t n = (B - r0 < rows_per) ? B - r0 : rows_per;
        cudaStreamWaitEvent(ctx->stream, ev[c], 0);
        LayerShapes s{n, i, o, TOPS_F32};
        const float* Xc = (const float*)Xd->data + r0 * i; const float* dAc = (const float*)dAd->data + r0 * o;
        float* Ac = (float*)Ad->data + r0 * o; float* dZc = (float*)dZ->data + r0 * o; float* dXc = (float*)dXd->data + r0 * i;
        if (f16x3) {
            rc = layer_fwd_grad_f16x3(ctx, s, Xc, wp, b ? (const float*)b->data : nullptr, act, dAc, Ac, dXc, dW, db, true, nullptr);
            continue;
        }
        int fused = 0;
        rc = fwd_gemm(ctx, s, Xc, W->data, b ? (const float*)b->data : nullptr, act, EPI_BIAS_ACT_DZ, Ac, dAc, dZc, nullptr, db, &fused, true, &ws);
        if (rc == TOPS_OK) rc = dw_db(ctx, s, dZc, Xc, dW, db, fused != 0, true);
        if (rc == TOPS_OK) rc = dx_gemm(ctx, s, dZc, W->data, dXc, EPI_STORE, ACT_ID, nullptr, nullptr, nullptr, &ws);
    }
    for (auto e : ev) cudaEventDestroy(e);
    cudaEventDestroy(ready);
    if (rc != TOPS_OK) { cudaStreamSynchronize(ctx->copy_stream); return rc; }
    if (grads_host) {
        CUDA_TRY(ctx, cudaMemcpyAsync(grads_host, dW, sizeof(float) * (size_t)(o * i + o), cudaMemcpyDeviceToHost, ctx->stream));
        return tops_sync(ctx);
    }
    return TOPS_OK;
}

// Data-parallel forward + VJP with the gradient all-reduce FUSED into the dW GEMM (NVLS): `grads_mc` is the multicast alias of a
// packed [dW (o*i) || db (o)] buffer that every rank of the job has bound (torch symmetric memory).  Split-K partials of dW are
// summed locally (`grads_local`); the last partial to finish a 32-row region of a tile pushes the finished region once with
// multimem.red — the NVSwitch adds it into every rank's replica while the remaining tiles are still being computed.  db (o floats)
// is pushed by a tiny kernel right after the forward GEMM.  No separate all-reduce pass over dW exists.
// Protocol (caller): zero the symmetric replica, barrier (all replicas zeroed), this call, barrier (all reductions landed).
extern "C" int tops_fflayer_fwd_grad_mc(tops_ctx* ctx, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act, const tops_buf* dA,
                                        tops_buf** A, tops_buf** dX, tops_buf** grads_local, void* grads_mc) {
    CHECK_CTX(ctx); LOCK(ctx);
    LayerShapes s; TRY(layer_shapes(ctx, X, W, b, &s)); TRY(check_act(ctx, act));
    if (!grads_mc || (reinterpret_cast<uintptr_t>(grads_mc) & 15) != 0) return set_err(ctx, TOPS_ERR_INVALID, "fflayer_fwd_grad_mc: multicast pointer must be non-NULL and 16-byte aligned");
    if (!dA || dA->rank != 2 || dA->dims[0] != s.B || dA->dims[1] != s.o || dA->dtype != s.dtype || dA->tr) return set_err(ctx, TOPS_ERR_SHAPE, "fflayer: dA[B,o] expected");
    if (!grads_local) return set_err(ctx, TOPS_ERR_INVALID, "fflayer_fwd_grad_mc: NULL grads_local slot");
    int64_t dAo[2] = {s.B, s.o}, dXs[2] = {s.B, s.i}, g_[1] = {s.o * s.i + s.o};
    TRY(prep_out(ctx, A, s.dtype, 2, dAo));
    if (dX) TRY(prep_out(ctx, dX, s.dtype, 2, dXs));
    TRY(prep_out(ctx, grads_local, TOPS_F32, 1, g_));
    float* dW = (float*)(*grads_local)->data; float* db = dW + s.o * s.i;
    float* dW_mc = (float*)grads_mc; float* db_mc = dW_mc + s.o * s.i;
    Tmp tmp;
    if (f16x3_layer_ok(ctx, s)) {
        WPairF16 wp; TRY(prep_wpair_f16(ctx, s, W->data, tmp, &wp, true));
        TRY(layer_fwd_grad_f16x3(ctx, s, X->data, wp, b ? (const float*)b->data : nullptr, act, dA->data, (*A)->data, dX ? (*dX)->data : nullptr, dW, db, false, dW_mc));
        k::mc_push(lc_of(ctx), db, db_mc, s.o);
        return check_launch(ctx, "mc_push");
    }
    tops_buf* dZ = nullptr;
    TRY(alloc_buf(ctx, s.dtype, 2, dAo, &dZ)); tmp.keep(dZ);
    WSplit ws; TRY(make_wsplit(ctx, W, tmp, &ws));
    int db_fused = 0;
    TRY(fwd_gemm(ctx, s, X->data, W->data, b ? (const float*)b->data : nullptr, act, EPI_BIAS_ACT_DZ, (*A)->data, dA->data, dZ->data, nullptr, db, &db_fused, false, &ws));
    TRY(dw_db(ctx, s, dZ->data, X->data, dW, db, db_fused != 0, false, dW_mc));
    k::mc_push(lc_of(ctx), db, db_mc, s.o);            // o floats: the bias gradient joins the same multicast buffer
    TRY(check_launch(ctx, "mc_push"));
    if (dX) TRY(dx_gemm(ctx, s, dZ->data, W->data, (*dX)->data, EPI_STORE, ACT_ID, nullptr, nullptr, nullptr, &ws));
    return TOPS_OK;
}

// ================================================================================================ recorded graphs
// The deferred evaluator SURVEY 8-b sketches (`tops_graph_begin/op/end/run`): instead of a separate op vocabulary, the API calls
// themselves are the ops.  Between begin and end every call is recorded (CUDA stream capture on the context's stream) instead of
// executed; tops_graph_launch replays the whole sequence — a composed TOp's forward and reverse sweep, an SGD step, a per-sample
// training step — as ONE cudaGraphLaunch: no per-method launch latency, no host round-trip between the Category-composed stages.
// Inputs are read from, and results written to, the same tensors on every replay (update inputs in place with tops_upload /
// tops_copy before launching).
extern "C" int tops_graph_begin(tops_ctx* ctx, size_t arena_bytes, tops_graph** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!out) return set_err(ctx, TOPS_ERR_INVALID, "tops_graph_begin: NULL slot");
    if (ctx->capturing) return set_err(ctx, TOPS_ERR_INVALID, "tops_graph_begin: already recording");
    if (ctx->profiling) return set_err(ctx, TOPS_ERR_INVALID, "tops_graph_begin: switch per-kernel profiling off first");
    tops_graph* g = new tops_graph();
    g->ctx = ctx; g->arena_bytes = arena_bytes ? arena_bytes : ((size_t)64 << 20);
    if (cudaMalloc((void**)&g->arena, g->arena_bytes) != cudaSuccess) { cudaGetLastError(); delete g; return set_err(ctx, TOPS_ERR_OOM, "graph arena of %zu bytes", arena_bytes); }
    cudaError_t e = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaFree(g->arena); delete g; return set_err(ctx, TOPS_ERR_CUDA, "cudaStreamBeginCapture: %s", cudaGetErrorString(e)); }
    g->launches_recorded = ctx->launches;
    ctx->capturing = g;
    *out = g;
    return TOPS_OK;
}
extern "C" int tops_graph_end(tops_ctx* ctx, tops_graph* g) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!g || ctx->capturing != g) return set_err(ctx, TOPS_ERR_INVALID, "tops_graph_end: not the graph being recorded");
    ctx->capturing = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g->graph);
    if (e != cudaSuccess || !g->graph) { cudaGetLastError(); return set_err(ctx, TOPS_ERR_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e)); }
    e = cudaGraphInstantiate(&g->exec, g->graph, 0);
    if (e != cudaSuccess) return set_err(ctx, TOPS_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    g->launches_recorded = ctx->launches - g->launches_recorded;
    return TOPS_OK;
}
extern "C" int tops_graph_launch(tops_ctx* ctx, tops_graph* g) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!g || !g->exec) return set_err(ctx, TOPS_ERR_INVALID, "tops_graph_launch: graph not finished");
    CUDA_TRY(ctx, cudaGraphLaunch(g->exec, ctx->stream));
    ctx->launches += g->launches_recorded;
    return TOPS_OK;
}
extern "C" int64_t tops_graph_kernel_count(const tops_graph* g) { return g ? g->launches_recorded : -1; }
extern "C" int tops_graph_destroy(tops_ctx* ctx, tops_graph* g) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!g) return TOPS_OK;
    if (ctx->capturing == g) { cudaGraph_t tmp = nullptr; cudaStreamEndCapture(ctx->stream, &tmp); if (tmp) cudaGraphDestroy(tmp); ctx->capturing = nullptr; cudaGetLastError(); }
    cudaStreamSynchronize(ctx->stream);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    if (g->arena) cudaFree(g->arena);
    delete g;
    return TOPS_OK;
}
// dst <- src (same element count and dtype), device to device: how a recorded step publishes its new state (e.g. the SGD-updated
// parameters) into the tensors its next replay reads
extern "C" int tops_copy(tops_ctx* ctx, tops_buf* dst, const tops_buf* src) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!dst || !src) return set_err(ctx, TOPS_ERR_INVALID, "tops_copy: NULL tensor");
    if (dst->tr || src->tr) return set_err(ctx, TOPS_ERR_INVALID, "tops_copy: transposed views are not accepted");
    if (dst->dtype != src->dtype || dst->numel != src->numel) return set_err(ctx, TOPS_ERR_SHAPE, "tops_copy: %lld elements of dtype %d into %lld of dtype %d", (long long)src->numel, src->dtype, (long long)dst->numel, dst->dtype);
    if (src->numel) CUDA_TRY(ctx, cudaMemcpyAsync(dst->data, src->data, (size_t)src->numel * esize(src->dtype), cudaMemcpyDeviceToDevice, ctx->stream));
    return TOPS_OK;
}

// ---- events (for schedules that span streams: the data-parallel step below)
extern "C" int tops_event_create(tops_ctx* ctx, void** ev) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!ev) return set_err(ctx, TOPS_ERR_INVALID, "tops_event_create: NULL slot");
    cudaEvent_t e;
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    *ev = e;
    return TOPS_OK;
}
extern "C" int tops_event_destroy(tops_ctx* ctx, void* ev) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (ev) CUDA_TRY(ctx, cudaEventDestroy((cudaEvent_t)ev));
    return TOPS_OK;
}
extern "C" int tops_stream_wait_event(tops_ctx* ctx, void* stream, void* ev) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!ev) return set_err(ctx, TOPS_ERR_INVALID, "tops_stream_wait_event: NULL event");
    CUDA_TRY(ctx, cudaStreamWaitEvent(stream ? (cudaStream_t)stream : ctx->stream, (cudaEvent_t)ev, 0));
    return TOPS_OK;
}

// Data-parallel step, library-owned schedule: forward, dW (+ db) into the packed buffer `grads` = [dW (o*i) || db (o)], then the
// event `grads_ready` is recorded and ONLY THEN the dX GEMM is launched — dX does not depend on dW, so the caller's all-reduce of
// `grads` (NCCL on a communication stream that waits for the event) runs concurrently with it.  `reserve_sms` SMs are left free by
// the persistent dX GEMM so that the collective's CTAs are scheduled at once instead of behind it.
extern "C" int tops_fflayer_step_dp(tops_ctx* ctx, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act, const tops_buf* dA,
                                    tops_buf** A, tops_buf** dX, tops_buf** grads, void* grads_ready, int reserve_sms) {
    CHECK_CTX(ctx); LOCK(ctx);
    LayerShapes s; TRY(layer_shapes(ctx, X, W, b, &s)); TRY(check_act(ctx, act));
    if (!dA || dA->rank != 2 || dA->dims[0] != s.B || dA->dims[1] != s.o || dA->dtype != s.dtype || dA->tr) return set_err(ctx, TOPS_ERR_SHAPE, "fflayer: dA[B,o] expected");
    if (!grads) return set_err(ctx, TOPS_ERR_INVALID, "fflayer_step_dp: NULL grads slot");
    if (reserve_sms < 0 || reserve_sms >= ctx->num_sms) return set_err(ctx, TOPS_ERR_INVALID, "fflayer_step_dp: reserve_sms out of range");
    int64_t dAo[2] = {s.B, s.o}, dXs[2] = {s.B, s.i}, g_[1] = {s.o * s.i + s.o};
    TRY(prep_out(ctx, A, s.dtype, 2, dAo));
    if (dX) TRY(prep_out(ctx, dX, s.dtype, 2, dXs));
    TRY(prep_out(ctx, grads, TOPS_F32, 1, g_));
    float* dW = (float*)(*grads)->data; float* db = dW + s.o * s.i;
    const int dx_ctas = reserve_sms > 0 ? ((ctx->num_sms - reserve_sms) & ~1) : 0;
    cudaEvent_t ev = (cudaEvent_t)grads_ready;
    Tmp tmp;
    if (s.B == 0) {
        CUDA_TRY(ctx, cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)(s.o * s.i + s.o), ctx->stream));
        if (ev) CUDA_TRY(ctx, cudaEventRecord(ev, ctx->stream));
        return TOPS_OK;
    }
    if (f16x3_layer_ok(ctx, s)) {
        WPairF16 wp; TRY(prep_wpair_f16(ctx, s, W->data, tmp, &wp, true));
        return layer_fwd_grad_f16x3(ctx, s, X->data, wp, b ? (const float*)b->data : nullptr, act, dA->data, (*A)->data, dX ? (*dX)->data : nullptr,
                                    dW, db, false, nullptr, ev, dx_ctas);
    }
    SplitScope split_scope_(ctx);
    tops_buf* dZ = nullptr;
    TRY(alloc_buf(ctx, s.dtype, 2, dAo, &dZ)); tmp.keep(dZ);
    WSplit ws; TRY(make_wsplit(ctx, W, tmp, &ws));
    int db_fused = 0;
    TRY(fwd_gemm(ctx, s, X->data, W->data, b ? (const float*)b->data : nullptr, act, EPI_BIAS_ACT_DZ, (*A)->data, dA->data, dZ->data, nullptr, db, &db_fused, false, &ws));
    TRY(dw_db(ctx, s, dZ->data, X->data, dW, db, db_fused != 0));
    if (ev) CUDA_TRY(ctx, cudaEventRecord(ev, ctx->stream));
    if (dX) TRY(dx_gemm(ctx, s, dZ->data, W->data, (*dX)->data, EPI_STORE, ACT_ID, nullptr, nullptr, nullptr, &ws, dx_ctas));
    return TOPS_OK;
}

extern "C" int tops_fflayer_grad(tops_ctx* ctx, const tops_buf* X, const tops_buf* W, const tops_buf* b, int act,
                                 const tops_buf* dA, const tops_buf* A_saved, tops_buf** dX, tops_buf** dW, tops_buf** db) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (!A_saved) {   // gradTOp' semantics: the forward is recomputed inside the gradient (Types.hs:155)
        tops_buf* A = nullptr;
        int r = tops_fflayer_fwd_grad(ctx, X, W, b, act, dA, &A, dX, dW, db);
        if (A) release_buf(A);
        return r;
    }
    LayerShapes s; TRY(layer_shapes(ctx, X, W, b, &s)); TRY(check_act(ctx, act));
    if (s.dtype != TOPS_F32) return set_err(ctx, TOPS_ERR_UNSUPPORTED, "fflayer_grad with a saved activation is fp32 only");
    if (!dA || dA->rank != 2 || dA->dims[0] != s.B || dA->dims[1] != s.o || !same_shape(dA, A_saved) || dA->tr || A_saved->tr) return set_err(ctx, TOPS_ERR_SHAPE, "fflayer: dA[B,o], A[B,o] expected");
    int64_t dAo[2] = {s.B, s.o}, dXs[2] = {s.B, s.i}, dWs[2] = {s.o, s.i}, dbs[1] = {s.o};
    if (dX) TRY(prep_out(ctx, dX, TOPS_F32, 2, dXs));
    TRY(prep_out(ctx, dW, TOPS_F32, 2, dWs));
    if (db) TRY(prep_out(ctx, db, TOPS_F32, 1, dbs));
    Tmp tmp; tops_buf* dZ = nullptr;
    SplitScope split_scope_(ctx);
    TRY(alloc_buf(ctx, TOPS_F32, 2, dAo, &dZ)); tmp.keep(dZ);
    k::dact_mul(lc_of(ctx), act, (const float*)dA->data, (const float*)A_saved->data, (float*)dZ->data, dZ->numel);
    TRY(check_launch(ctx, "dact_mul"));
    TRY(dw_db(ctx, s, dZ->data, X->data, (float*)(*dW)->data, db ? (float*)(*db)->data : nullptr));
    if (dX) TRY(dx_gemm(ctx, s, dZ->data, W->data, (*dX)->data, EPI_STORE, ACT_ID, nullptr));
    return TOPS_OK;
}

namespace {

int mlp_check(tops_ctx* ctx, int n, const tops_buf* const* W, const tops_buf* const* b, const int* acts, const tops_buf* X) {
    if (n < 1 || !W || !b || !acts || !X) return set_err(ctx, TOPS_ERR_INVALID, "mlp: NULL argument");
    if (X->rank != 2 || X->dtype != TOPS_F32 || X->tr) return set_err(ctx, TOPS_ERR_SHAPE, "mlp: X[B,i] fp32 expected");
    int64_t in = X->dims[1];
    for (int l = 0; l < n; ++l) {
        if (!W[l] || !b[l] || W[l]->dtype != TOPS_F32 || b[l]->dtype != TOPS_F32 || W[l]->tr || W[l]->rank != 2 || b[l]->rank != 1 || W[l]->dims[1] != in || b[l]->dims[0] != W[l]->dims[0])
            return set_err(ctx, TOPS_ERR_SHAPE, "mlp: layer %d: W[o,i] b[o] do not chain (i=%lld)", l, (long long)in);
        if (acts[l] < TOPS_ACT_ID || acts[l] > TOPS_ACT_SOFTMAX) return set_err(ctx, TOPS_ERR_INVALID, "mlp: unknown activation %d", acts[l]);
        in = W[l]->dims[0];
    }
    return TOPS_OK;
}

// one forward layer: returns Z (only for softmax, else NULL) and A
int mlp_layer_fwd(tops_ctx* ctx, const tops_buf* in, const tops_buf* W, const tops_buf* b, int act, Tmp& tmp, tops_buf** Zout, tops_buf* A,
                  const WSplit* ws = nullptr) {
    LayerShapes s{in->dims[0], in->dims[1], W->dims[0], TOPS_F32};
    *Zout = nullptr;
    if (act == TOPS_ACT_SOFTMAX) {
        int64_t d[2] = {s.B, s.o};
        tops_buf* Z = nullptr;
        TRY(alloc_buf(ctx, TOPS_F32, 2, d, &Z)); tmp.keep(Z);
        TRY(fwd_gemm(ctx, s, in->data, W->data, (const float*)b->data, ACT_ID, EPI_BIAS_ACT, Z->data, nullptr, nullptr, nullptr, nullptr, nullptr, false, ws));
        if (A) { k::softmax_rows(lc_of(ctx), (const float*)Z->data, (float*)A->data, s.B, s.o); TRY(check_launch(ctx, "softmax")); }
        *Zout = Z;
        return TOPS_OK;
    }
    return fwd_gemm(ctx, s, in->data, W->data, (const float*)b->data, act, EPI_BIAS_ACT, A->data, nullptr, nullptr, nullptr, nullptr, nullptr, false, ws);
}

}  // namespace

extern "C" int tops_mlp_fwd(tops_ctx* ctx, int n, const tops_buf* const* W, const tops_buf* const* b, const int* acts, const tops_buf* X, tops_buf** A_out) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(mlp_check(ctx, n, W, b, acts, X));
    Tmp tmp;
    SplitScope split_scope_(ctx);   // F16X3: every tensor is split into its fp16 pair at most once per call
    const tops_buf* cur = X;
    for (int l = 0; l < n; ++l) {
        int64_t d[2] = {X->dims[0], W[l]->dims[0]};
        tops_buf* A = nullptr;
        if (l == n - 1) { TRY(prep_out(ctx, A_out, TOPS_F32, 2, d)); A = *A_out; }
        else { TRY(alloc_buf(ctx, TOPS_F32, 2, d, &A)); tmp.keep(A); }
        tops_buf* Z;
        WSplit ws; TRY(make_wsplit(ctx, W[l], tmp, &ws));
        TRY(mlp_layer_fwd(ctx, cur, W[l], b[l], acts[l], tmp, &Z, A, &ws));
        cur = A;
    }
    return TOPS_OK;
}

extern "C" int tops_mlp_fwd_grad(tops_ctx* ctx, int n, const tops_buf* const* W, const tops_buf* const* b, const int* acts, int loss,
                                 const tops_buf* X, const tops_buf* Y, tops_buf** A_out, tops_buf** loss_sum, tops_buf** dX,
                                 tops_buf** dW, tops_buf** db) {
    CHECK_CTX(ctx); LOCK(ctx);
    TRY(mlp_check(ctx, n, W, b, acts, X));
    if (loss != TOPS_LOSS_SQUARED_ERROR && loss != TOPS_LOSS_CROSS_ENTROPY) return set_err(ctx, TOPS_ERR_INVALID, "mlp: unknown loss %d", loss);
    const int64_t B = X->dims[0], o_last = W[n - 1]->dims[0];
    if (!Y || Y->dtype != TOPS_F32 || Y->rank != 2 || Y->dims[0] != B || Y->dims[1] != o_last || Y->tr) return set_err(ctx, TOPS_ERR_SHAPE, "mlp: Y[B,o] fp32 expected");
    if (!dW || !db || !A_out || !loss_sum) return set_err(ctx, TOPS_ERR_INVALID, "mlp: NULL output slot");
    Tmp tmp;
    SplitScope split_scope_(ctx);   // F16X3: every tensor (activations, dZ, W) is split into its fp16 pair at most once per call
    split_scope_.track_max = true;
    std::vector<const tops_buf*> acts_in(n);     // input of layer l
    std::vector<tops_buf*> Zs(n, nullptr);
    TRY(prep_out(ctx, loss_sum, TOPS_F32, 0, nullptr));
    float* lossp = (float*)(*loss_sum)->data;
    CUDA_TRY(ctx, cudaMemsetAsync(lossp, 0, 4, ctx->stream));
    std::vector<WSplit> wsp(n);   // bf16 correction operands of every W: used by its forward GEMM and its dX GEMM
    for (int l = 0; l < n; ++l) TRY(make_wsplit(ctx, W[l], tmp, &wsp[l]));
    // gradient outputs first: the GEMM epilogues that produce a dZ also produce that layer's db (fused column sums)
    std::vector<int> db_done(n, 0);
    for (int l = 0; l < n; ++l) {
        int64_t dWs[2] = {W[l]->dims[0], W[l]->dims[1]}, dbs[1] = {W[l]->dims[0]};
        TRY(prep_out(ctx, &dW[l], TOPS_F32, 2, dWs));
        TRY(prep_out(ctx, &db[l], TOPS_F32, 1, dbs));
    }
    // ---- forward; the last layer's epilogue also produces the loss and dZ when the pairing allows
    const tops_buf* cur = X;
    tops_buf* dZ = nullptr;
    for (int l = 0; l < n; ++l) {
        acts_in[l] = cur;
        int64_t d[2] = {B, W[l]->dims[0]};
        tops_buf* A = nullptr;
        const bool last = (l == n - 1);
        if (last) { TRY(prep_out(ctx, A_out, TOPS_F32, 2, d)); A = *A_out; }
        else { TRY(alloc_buf(ctx, TOPS_F32, 2, d, &A)); tmp.keep(A); }
        if (last) { TRY(alloc_buf(ctx, TOPS_F32, 2, d, &dZ)); tmp.keep(dZ); }
        if (last && acts[l] != TOPS_ACT_SOFTMAX && loss == TOPS_LOSS_SQUARED_ERROR) {
            LayerShapes s{B, cur->dims[1], d[1], TOPS_F32};
            TRY(fwd_gemm(ctx, s, cur->data, W[l]->data, (const float*)b[l]->data, acts[l], EPI_BIAS_ACT_SE, A->data, Y->data, dZ->data, lossp,
                         (float*)db[l]->data, &db_done[l], false, &wsp[l]));
        } else if (last && acts[l] == TOPS_ACT_SOFTMAX && loss == TOPS_LOSS_CROSS_ENTROPY) {
            TRY(mlp_layer_fwd(ctx, cur, W[l], b[l], acts[l], tmp, &Zs[l], nullptr, &wsp[l]));
            CUDA_TRY(ctx, cudaMemsetAsync(db[l]->data, 0, sizeof(float) * (size_t)d[1], ctx->stream));
            db_done[l] = k::softmax_ce_rows(lc_of(ctx), (const float*)Zs[l]->data, (const float*)Y->data, (float*)A->data, (float*)dZ->data, lossp, B, d[1],
                                            (float*)db[l]->data) ? 1 : 0;   // narrow heads: db comes out of the same pass
            TRY(check_launch(ctx, "softmax_ce"));
        } else {
            TRY(mlp_layer_fwd(ctx, cur, W[l], b[l], acts[l], tmp, &Zs[l], A, &wsp[l]));
            if (last) {   // generic head: loss VJP on activations, then the activation's VJP
                tops_buf* dA = nullptr;
                TRY(alloc_buf(ctx, TOPS_F32, 2, d, &dA)); tmp.keep(dA);
                k::loss_vjp(lc_of(ctx), loss, (const float*)A->data, (const float*)Y->data, (float*)dA->data, lossp, A->numel);
                if (acts[l] == TOPS_ACT_SOFTMAX) k::softmax_vjp_rows(lc_of(ctx), (const float*)Zs[l]->data, (const float*)dA->data, (float*)dZ->data, B, d[1]);
                else k::dact_mul(lc_of(ctx), acts[l], (const float*)dA->data, (const float*)A->data, (float*)dZ->data, A->numel);
                TRY(check_launch(ctx, "loss head"));
            }
        }
        cur = A;
    }
    // ---- reverse sweep: dW_l = dZ_l^T in_l, db_l = Σ dZ_l, dZ_{l-1} = (dZ_l W_l) ⊙ act'(A_{l-1}) fused in the GEMM epilogue
    for (int l = n - 1; l >= 0; --l) {
        LayerShapes s{B, W[l]->dims[1], W[l]->dims[0], TOPS_F32};
        TRY(dw_db(ctx, s, dZ->data, acts_in[l]->data, (float*)dW[l]->data, (float*)db[l]->data, db_done[l] != 0));
        if (l > 0) {
            int64_t d[2] = {B, s.i};
            tops_buf* dZp = nullptr;
            TRY(alloc_buf(ctx, TOPS_F32, 2, d, &dZp)); tmp.keep(dZp);
            if (acts[l - 1] == TOPS_ACT_SOFTMAX) {
                tops_buf* dAp = nullptr;
                TRY(alloc_buf(ctx, TOPS_F32, 2, d, &dAp)); tmp.keep(dAp);
                TRY(dx_gemm(ctx, s, dZ->data, W[l]->data, dAp->data, EPI_STORE, ACT_ID, nullptr, nullptr, nullptr, &wsp[l]));
                k::softmax_vjp_rows(lc_of(ctx), (const float*)Zs[l - 1]->data, (const float*)dAp->data, (float*)dZp->data, B, s.i);
                TRY(check_launch(ctx, "softmax_vjp"));
            } else {
                TRY(dx_gemm(ctx, s, dZ->data, W[l]->data, dZp->data, EPI_MUL_DACT, acts[l - 1], acts_in[l]->data,
                            (float*)db[l - 1]->data, &db_done[l - 1], &wsp[l]));
            }
            dZ = dZp;
        } else if (dX) {
            int64_t d[2] = {B, s.i};
            TRY(prep_out(ctx, dX, TOPS_F32, 2, d));
            TRY(dx_gemm(ctx, s, dZ->data, W[l]->data, (*dX)->data, EPI_STORE, ACT_ID, nullptr, nullptr, nullptr, &wsp[l]));
        }
    }
    return TOPS_OK;
}

extern "C" int tops_sgd_step(tops_ctx* ctx, int n, const tops_buf* const* params, const tops_buf* const* grads, double rate, tops_buf** out) {
    CHECK_CTX(ctx); LOCK(ctx);
    if (n < 0 || !params || !grads || !out) return set_err(ctx, TOPS_ERR_INVALID, "sgd_step: NULL argument");
    Tmp tmp;
    for (int j = 0; j < n; ++j) {
        TRY(need_f32(ctx, params[j], "tops_sgd_step")); TRY(need_f32(ctx, grads[j], "tops_sgd_step"));
        if (!same_shape(params[j], grads[j])) return set_err(ctx, TOPS_ERR_SHAPE, "sgd_step: parameter %d and its gradient differ in shape", j);
        const tops_buf *ps, *gs;     // O(1) transposed views (e.g. a gradient that came through transpOp) are materialised
        TRY(contig(ctx, params[j], tmp, &ps)); TRY(contig(ctx, grads[j], tmp, &gs));
        TRY(prep_out(ctx, &out[j], TOPS_F32, ps->rank, ps->dims));
        k::sgd(lc_of(ctx), (const float*)ps->data, (const float*)gs->data, (float)rate, (float*)out[j]->data, ps->numel);
    }
    return check_launch(ctx, "sgd_step");
}
