// Bandwidth-bound kernels of libtops_b200: generation, elementwise (incl. the `liftT` bytecode interpreter),
// deterministic two-stage reductions, BLAS-2, layout permutation, the softmax / loss heads, and a CUDA-core GEMM
// that honours the same operand/epilogue contract as the tcgen05 engine.  Grid-stride loops sized in multiples
// of the SM count, 128-bit accesses where alignment allows, warp-shuffle reductions.
#include "kernels.h"

#include <cuda_bf16.h>

#include <initializer_list>

#include "gemm_sm100.cuh"   // epilogue enums, act_apply / act_deriv_from_out

namespace tops {
namespace k {

namespace {

constexpr int kThreads = 256;

inline int grid_for(const LaunchCtx& lc, int64_t work_items, int per_block = kThreads, int waves = 8) {
    int64_t b = (work_items + per_block - 1) / per_block;
    int64_t cap = (int64_t)lc.num_sms * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
inline void count(const LaunchCtx& lc) { if (lc.launches) ++*lc.launches; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
// 2-D grid for row/column kernels: x covers `col_items` work items (kThreads per block), y walks the rows with ~8 waves of CTAs in total
inline dim3 grid_rows_cols(const LaunchCtx& lc, int64_t rows, int64_t col_items) {
    const int64_t gx = (col_items + kThreads - 1) / kThreads;
    int64_t gy = ((int64_t)lc.num_sms * 8 + gx - 1) / gx;
    if (gy > rows) gy = rows;
    if (gy > 65535) gy = 65535;
    if (gy < 1) gy = 1;
    return dim3((unsigned)gx, (unsigned)gy);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float block_sum(float v, float* sh) {   // sh: >= 32 floats; result valid in thread 0
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) sh[w] = v;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
        r = (l < (blockDim.x >> 5)) ? sh[l] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------ Philox4x32-10
__device__ __forceinline__ uint4 philox(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ float u01(uint32_t x) { return (x >> 8) * (1.0f / 16777216.0f) + (0.5f / 16777216.0f); }

// 16-byte stores when the base is 16-byte aligned, scalar tail
__global__ void k_fill(float* p, int64_t n, float v, bool vec) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
    const float4 v4 = make_float4(v, v, v, v);
    for (int64_t i = tid; i < n4; i += nth) reinterpret_cast<float4*>(p)[i] = v4;
    for (int64_t i = n4 * 4 + tid; i < n; i += nth) p[i] = v;
}
__global__ void k_fill_bf16(__nv_bfloat16* p, int64_t n, float v) {
    const __nv_bfloat16 b = __float2bfloat16_rn(v);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = b;
}
template <bool NORMAL>
__global__ void k_rand(float* p, int64_t n, float a, float b, uint64_t seed) {
    const int64_t n4 = (n + 3) / 4;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        uint4 r = philox(make_uint4((uint32_t)i, (uint32_t)(i >> 32), 0u, 0u), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
        float v[4];
        if (NORMAL) {   // Box-Muller, mean a, sd b
            const float r0 = sqrtf(-2.0f * logf(u01(r.x))), r1 = sqrtf(-2.0f * logf(u01(r.z)));
            float s0, c0, s1, c1;
            sincospif(2.0f * u01(r.y), &s0, &c0);
            sincospif(2.0f * u01(r.w), &s1, &c1);
            v[0] = a + b * r0 * c0; v[1] = a + b * r0 * s0; v[2] = a + b * r1 * c1; v[3] = a + b * r1 * s1;
        } else {        // uniform [a, b)
            v[0] = a + (b - a) * u01(r.x); v[1] = a + (b - a) * u01(r.y); v[2] = a + (b - a) * u01(r.z); v[3] = a + (b - a) * u01(r.w);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) if (i * 4 + e < n) p[i * 4 + e] = v[e];
    }
}
// 8 elements per thread: 2 x 16-byte fp32 accesses against one 16-byte bf16 access
__global__ void k_cast_f2b(const float* __restrict__ s, __nv_bfloat16* __restrict__ d, int64_t n, bool vec) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n8 = vec ? n / 8 : 0;
    for (int64_t i = tid; i < n8; i += nth) {
        const float4 a = reinterpret_cast<const float4*>(s)[2 * i], b = reinterpret_cast<const float4*>(s)[2 * i + 1];
        uint4 u;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
        h[0] = __floats2bfloat162_rn(a.x, a.y); h[1] = __floats2bfloat162_rn(a.z, a.w);
        h[2] = __floats2bfloat162_rn(b.x, b.y); h[3] = __floats2bfloat162_rn(b.z, b.w);
        reinterpret_cast<uint4*>(d)[i] = u;
    }
    for (int64_t i = n8 * 8 + tid; i < n; i += nth) d[i] = __float2bfloat16_rn(s[i]);
}
__global__ void k_cast_b2f(const __nv_bfloat16* __restrict__ s, float* __restrict__ d, int64_t n, bool vec) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n8 = vec ? n / 8 : 0;
    for (int64_t i = tid; i < n8; i += nth) {
        const uint4 u = reinterpret_cast<const uint4*>(s)[i];
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
        const float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]), f2 = __bfloat1622float2(h[2]), f3 = __bfloat1622float2(h[3]);
        reinterpret_cast<float4*>(d)[2 * i] = make_float4(f0.x, f0.y, f1.x, f1.y);
        reinterpret_cast<float4*>(d)[2 * i + 1] = make_float4(f2.x, f2.y, f3.x, f3.y);
    }
    for (int64_t i = n8 * 8 + tid; i < n; i += nth) d[i] = __bfloat162float(s[i]);
}
__global__ void k_eye(float* p, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n * n; i += (int64_t)gridDim.x * blockDim.x) p[i] = (i / n == i % n) ? 1.f : 0.f;
}

// ------------------------------------------------------------------ elementwise, float4 body + scalar tail
template <typename F>
__global__ void k_map1(const float* __restrict__ x, float* __restrict__ out, int64_t n, bool vec, F f) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
    for (int64_t i = tid; i < n4; i += nth) {
        float4 a = reinterpret_cast<const float4*>(x)[i];
        reinterpret_cast<float4*>(out)[i] = make_float4(f(a.x), f(a.y), f(a.z), f(a.w));
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += nth) out[i] = f(x[i]);
}
template <typename F>
__global__ void k_map2(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int64_t n, bool vec, F f) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
    for (int64_t i = tid; i < n4; i += nth) {
        float4 a = reinterpret_cast<const float4*>(x)[i], b = reinterpret_cast<const float4*>(y)[i];
        reinterpret_cast<float4*>(out)[i] = make_float4(f(a.x, b.x), f(a.y, b.y), f(a.z, b.z), f(a.w, b.w));
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += nth) out[i] = f(x[i], y[i]);
}

struct InPtrs { const float* p[8]; };

__global__ void k_add_n(InPtrs in, int n_in, float* __restrict__ out, int64_t n, bool vec) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t n4 = vec ? n / 4 : 0;
    for (int64_t i = tid; i < n4; i += nth) {
        float4 acc = reinterpret_cast<const float4*>(in.p[0])[i];
        for (int j = 1; j < n_in; ++j) {
            float4 b = reinterpret_cast<const float4*>(in.p[j])[i];
            acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
        }
        reinterpret_cast<float4*>(out)[i] = acc;
    }
    for (int64_t i = n4 * 4 + tid; i < n; i += nth) {
        float acc = in.p[0][i];
        for (int j = 1; j < n_in; ++j) acc += in.p[j][i];
        out[i] = acc;
    }
}

// [rows, cols] row-major, no div/mod per element: blockIdx.x / threadIdx.x walk 4-column chunks (the thread's bias chunk stays in
// registers), blockIdx.y walks rows.  VEC: cols % 4 == 0 and 16-byte aligned bases.
template <bool VEC>
__global__ void k_bias_act(int act, const float* __restrict__ Z, const float* __restrict__ bias, float* __restrict__ A, int64_t rows, int64_t cols) {
    constexpr int W = VEC ? 4 : 1;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * W;
    if (c >= cols) return;
    float b[W];
#pragma unroll
    for (int e = 0; e < W; ++e) b[e] = bias ? bias[c + e] : 0.f;
    for (int64_t r = blockIdx.y; r < rows; r += gridDim.y) {
        if constexpr (VEC) {
            const float4 z = *reinterpret_cast<const float4*>(Z + r * cols + c);
            *reinterpret_cast<float4*>(A + r * cols + c) = make_float4(act_apply(act, z.x + b[0]), act_apply(act, z.y + b[1]), act_apply(act, z.z + b[2]), act_apply(act, z.w + b[3]));
        } else {
            A[r * cols + c] = act_apply(act, Z[r * cols + c] + b[0]);
        }
    }
}

// ------------------------------------------------------------------ liftT interpreter: 4 elements per thread
__device__ __forceinline__ float lift_unary(int op, float a) {
    switch (op) {
        case 6: return -a;
        case 7: return __expf(a);
        case 8: return __logf(a);
        case 9: return __fdividef(1.0f, a);
        case 10: return sqrtf(a);
        case 11: return tanhf(a);
        case 12: return fabsf(a);
        case 13: return (a > 0.f) - (a < 0.f);
        case 17: return __fdividef(1.0f, 1.0f + __expf(-a));
        case 18: return sinf(a);
        case 19: return cosf(a);
    }
    return a;
}
__device__ __forceinline__ float lift_binary(int op, float a, float b) {
    switch (op) {
        case 2: return a + b;
        case 3: return a - b;
        case 4: return a * b;
        case 5: return a / b;
        case 14: return fmaxf(a, b);
        case 15: return fminf(a, b);
        case 16: return powf(a, b);
    }
    return a;
}
__global__ void k_lift(LiftProgram prog, InPtrs in, int n_in, float* __restrict__ out, int64_t n, bool vec) {
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const int64_t ngroups = (n + 3) / 4;
    for (int64_t g = tid; g < ngroups; g += nth) {
        const int64_t base = g * 4;
        const int cnt = (int)min((int64_t)4, n - base);
        float x[8][4];
        for (int j = 0; j < n_in; ++j) {
            if (vec && cnt == 4) {
                float4 v = reinterpret_cast<const float4*>(in.p[j])[g];
                x[j][0] = v.x; x[j][1] = v.y; x[j][2] = v.z; x[j][3] = v.w;
            } else {
                for (int e = 0; e < 4; ++e) x[j][e] = e < cnt ? in.p[j][base + e] : 1.0f;
            }
        }
        float st[16][4];
        int sp = 0;
        for (int pc = 0; pc < prog.len; ++pc) {
            const int op = prog.code[pc] >> 16, arg = prog.code[pc] & 0xffff;
            if (op == 0) {
#pragma unroll
                for (int e = 0; e < 4; ++e) st[sp][e] = x[arg][e];
                ++sp;
            } else if (op == 1) {
                const float c = prog.consts[arg];
#pragma unroll
                for (int e = 0; e < 4; ++e) st[sp][e] = c;
                ++sp;
            } else if (op == 2 || op == 3 || op == 4 || op == 5 || op == 14 || op == 15 || op == 16) {
                --sp;
#pragma unroll
                for (int e = 0; e < 4; ++e) st[sp - 1][e] = lift_binary(op, st[sp - 1][e], st[sp][e]);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) st[sp - 1][e] = lift_unary(op, st[sp - 1][e]);
            }
        }
        if (vec && cnt == 4) reinterpret_cast<float4*>(out)[g] = make_float4(st[0][0], st[0][1], st[0][2], st[0][3]);
        else for (int e = 0; e < cnt; ++e) out[base + e] = st[0][e];
    }
}

// ------------------------------------------------------------------ reductions
template <bool DOT>
__global__ void k_reduce_partial(const float* __restrict__ x, const float* __restrict__ y, int64_t n, float* __restrict__ partial) {
    __shared__ float sh[32];
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
    const bool vec = (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (!DOT || (reinterpret_cast<uintptr_t>(y) & 15) == 0);
    const int64_t n4 = vec ? n / 4 : 0;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains: 16-byte loads, two of them in flight per thread
    int64_t i = tid;
    for (; i + nth < n4; i += 2 * nth) {
        const float4 u = reinterpret_cast<const float4*>(x)[i], v = reinterpret_cast<const float4*>(x)[i + nth];
        if (DOT) {
            const float4 p = reinterpret_cast<const float4*>(y)[i], q = reinterpret_cast<const float4*>(y)[i + nth];
            a0 = fmaf(u.x, p.x, a0); a1 = fmaf(u.y, p.y, a1); a2 = fmaf(u.z, p.z, a2); a3 = fmaf(u.w, p.w, a3);
            a0 = fmaf(v.x, q.x, a0); a1 = fmaf(v.y, q.y, a1); a2 = fmaf(v.z, q.z, a2); a3 = fmaf(v.w, q.w, a3);
        } else {
            a0 += u.x + v.x; a1 += u.y + v.y; a2 += u.z + v.z; a3 += u.w + v.w;
        }
    }
    for (; i < n4; i += nth) {
        const float4 u = reinterpret_cast<const float4*>(x)[i];
        if (DOT) { const float4 p = reinterpret_cast<const float4*>(y)[i]; a0 = fmaf(u.x, p.x, a0); a1 = fmaf(u.y, p.y, a1); a2 = fmaf(u.z, p.z, a2); a3 = fmaf(u.w, p.w, a3); }
        else { a0 += u.x; a1 += u.y; a2 += u.z; a3 += u.w; }
    }
    for (int64_t j = n4 * 4 + tid; j < n; j += nth) a0 += DOT ? x[j] * y[j] : x[j];
    float acc = (a0 + a1) + (a2 + a3);
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = acc;
}
__global__ void k_reduce_final(const float* __restrict__ partial, int np, float* __restrict__ out) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < np; i += blockDim.x) acc += partial[i];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) *out = acc;
}
__global__ void k_trace(const float* __restrict__ a, int64_t n, int64_t ld, float* __restrict__ out) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) acc += a[i * ld + i];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) *out = acc;
}

// column sums / transposed gemv: partial[chunk][c] = sum_{r in chunk} w[r] * x[r, c]
template <typename TIn>
__global__ void k_colsum_partial(const TIn* __restrict__ x, const float* __restrict__ w, int64_t rows, int64_t cols, int64_t rows_per_chunk, float* __restrict__ partial) {
    __shared__ float sh[8][128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t c0 = (int64_t)blockIdx.x * 128;
    const int64_t r0 = (int64_t)blockIdx.y * rows_per_chunk;
    const int64_t r1 = min(rows, r0 + rows_per_chunk);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // fp32 with 16-byte aligned rows: lane owns 4 CONSECUTIVE columns (one 16-byte load per row); otherwise columns lane + 32 j
    const bool vec4 = sizeof(TIn) == 4 && (cols % 4) == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && c0 + 128 <= cols;
    if (vec4) {
        for (int64_t r = r0 + warp; r < r1; r += 8) {
            const float wr = w ? w[r] : 1.0f;
            const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + r * cols + c0 + lane * 4);
            acc[0] = fmaf(wr, v.x, acc[0]); acc[1] = fmaf(wr, v.y, acc[1]); acc[2] = fmaf(wr, v.z, acc[2]); acc[3] = fmaf(wr, v.w, acc[3]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[warp][lane * 4 + j] = acc[j];
    } else {
        for (int64_t r = r0 + warp; r < r1; r += 8) {
            const float wr = w ? w[r] : 1.0f;
            const TIn* row = x + r * cols + c0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t c = c0 + lane + 32 * j;
                if (c < cols) acc[j] = fmaf(wr, (float)row[lane + 32 * j], acc[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sh[warp][lane + 32 * j] = acc[j];
    }
    __syncthreads();
    if (threadIdx.x < 128) {
        float s = 0.f;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) s += sh[wv][threadIdx.x];
        const int64_t c = c0 + threadIdx.x;
        if (c < cols) partial[(int64_t)blockIdx.y * cols + c] = s;
    }
}
__global__ void k_colsum_final(const float* __restrict__ partial, int chunks, int64_t cols, float alpha, float beta, const float* y, float* out) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < cols; c += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k2 = 0; k2 < chunks; ++k2) s += partial[(int64_t)k2 * cols + c];
        out[c] = alpha * s + (y ? beta * y[c] : 0.f);
    }
}

// ------------------------------------------------------------------ BLAS-2 / layout
// out[r, c] = (x ? x[r] : 1) * y[c]: `ger` (x != NULL) and `broadcast_rows` (x == NULL).  Same 2-D scheme as k_bias_act: the
// thread's 4 columns of y stay in registers, rows are walked by blockIdx.y — one 16-byte store per row and thread, no div/mod.
template <bool VEC>
__global__ void k_outer_rows(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out, int64_t n, int64_t m) {
    constexpr int W = VEC ? 4 : 1;
    const int64_t c = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * W;
    if (c >= m) return;
    float v[W];
#pragma unroll
    for (int e = 0; e < W; ++e) v[e] = y[c + e];
    for (int64_t r = blockIdx.y; r < n; r += gridDim.y) {
        const float a = x ? x[r] : 1.0f;
        if constexpr (VEC) *reinterpret_cast<float4*>(out + r * m + c) = make_float4(a * v[0], a * v[1], a * v[2], a * v[3]);
        else out[r * m + c] = a * v[0];
    }
}
// one warp per output row; A row-major [n,m]
__global__ void k_gemv_rows(float alpha, const float* __restrict__ a, const float* __restrict__ x, float beta, const float* __restrict__ y, float* __restrict__ out, int64_t n, int64_t m, bool vec) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const float* row = a + r * m;
        float acc = 0.f;
        if (vec) {
            for (int64_t c = lane * 4; c < m; c += 128) {
                const float4 av = *reinterpret_cast<const float4*>(row + c), xv = *reinterpret_cast<const float4*>(x + c);
                acc += av.x * xv.x + av.y * xv.y + av.z * xv.z + av.w * xv.w;
            }
        } else {
            for (int64_t c = lane; c < m; c += 32) acc = fmaf(row[c], x[c], acc);
        }
        acc = warp_sum(acc);
        if (lane == 0) out[r] = alpha * acc + (y ? beta * y[r] : 0.f);
    }
}
__global__ void k_transpose(const float* __restrict__ in, float* __restrict__ out, int64_t rows, int64_t cols) {
    __shared__ float t[32][33];
    const int64_t bx = (int64_t)blockIdx.x * 32, by = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t r = by + j, c = bx + threadIdx.x;
        if (r < rows && c < cols) t[j][threadIdx.x] = in[r * cols + c];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        const int64_t c = bx + j, r = by + threadIdx.x;
        if (r < rows && c < cols) out[c * rows + r] = t[threadIdx.x][j];
    }
}
struct PermArgs { int rank; int64_t out_dims[8]; int64_t in_strides_for_out[8]; };
__global__ void k_permute(const float* __restrict__ in, float* __restrict__ out, PermArgs pa, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t rem = i, off = 0;
        for (int a = pa.rank - 1; a >= 0; --a) {
            const int64_t idx = rem % pa.out_dims[a];
            rem /= pa.out_dims[a];
            off += idx * pa.in_strides_for_out[a];
        }
        out[i] = in[off];
    }
}
// Permutation whose output's fastest axis is NOT the input's fastest axis (e.g. `transp` of a rank-3 tensor = full reversal):
// a 32 x 32 shared-memory tile over (p = the input's last axis, q = the axis that becomes the output's last) makes both the reads
// and the writes coalesced; the remaining axes form a batch whose index is decomposed ONCE per block, not per element.
struct PermTiled { int nb; int64_t bdim[8], bin[8], bout[8]; int64_t P, Q, p_out_stride, q_in_stride; };
__global__ void k_permute_tiled(const float* __restrict__ in, float* __restrict__ out, PermTiled pa) {
    __shared__ float t[32][33];
    int64_t rem = blockIdx.z, ioff = 0, ooff = 0;
    for (int a = pa.nb - 1; a >= 0; --a) {
        const int64_t idx = rem % pa.bdim[a];
        rem /= pa.bdim[a];
        ioff += idx * pa.bin[a]; ooff += idx * pa.bout[a];
    }
    const int64_t p0 = (int64_t)blockIdx.x * 32, q0 = (int64_t)blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {        // read: rows = q, contiguous along p
        const int64_t q = q0 + j, pp = p0 + threadIdx.x;
        if (q < pa.Q && pp < pa.P) t[j][threadIdx.x] = in[ioff + q * pa.q_in_stride + pp];
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {        // write: rows = p, contiguous along q
        const int64_t pp = p0 + j, q = q0 + threadIdx.x;
        if (q < pa.Q && pp < pa.P) out[ooff + pp * pa.p_out_stride + q] = t[threadIdx.x][j];
    }
}
__global__ void k_diag_embed(const float* __restrict__ v, float* __restrict__ out, int64_t n, int64_t step) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i * step] = v[i];
}
__global__ void k_diag_extract(const float* __restrict__ a, float* __restrict__ out, int64_t n, int64_t step) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) out[i] = a[i * step];
}

// ------------------------------------------------------------------ softmax / loss heads: one warp per row (sample)
// mode 0: A = softmax(Z)    mode 1: dZ = softmax-VJP(Z, dA)    mode 2: fused softmax + crossEntropy (A, loss, dZ)
template <int MODE>
__global__ void k_softmax_rows(const float* __restrict__ Z, const float* __restrict__ aux, float* __restrict__ A, float* __restrict__ dZ, float* __restrict__ loss, int64_t rows, int64_t cols,
                               float* __restrict__ colsum /* MODE 2, cols <= 32: += column sums of dZ (the head's db); NULL = none */) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float loss_acc = 0.f, col_acc = 0.f;               // lane c always handles column c when cols <= 32
    for (int64_t r = warp0; r < rows; r += nwarps) {
        const float* z = Z + r * cols;
        float se = 0.f;
        for (int64_t c = lane; c < cols; c += 32) se += __expf(z[c]);        // map exp >>> sumRows (no max-subtraction: NeuralNet.hs:52-59)
        se = warp_sum(se);
        const float rinv = 1.0f / se;                                          // map recip
        if (MODE == 0) {
            for (int64_t c = lane; c < cols; c += 32) A[r * cols + c] = __expf(z[c]) * rinv;   // outer LZ (LS LZ): scalar * vector
        } else {
            // dA: given (MODE 1) or from crossEntropy: dA = -(y / a)   (NeuralNet.hs:71-77 VJP)
            float s = 0.f;
            for (int64_t c = lane; c < cols; c += 32) {
                const float e = __expf(z[c]);
                float d;
                if (MODE == 2) {
                    const float a = e * rinv, yv = aux[r * cols + c];
                    A[r * cols + c] = a;
                    loss_acc -= __logf(a) * yv;
                    d = -(yv / a);
                } else d = aux[r * cols + c];
                s += d * e;
            }
            s = warp_sum(s);                                                   // VJP of outer wrt the scalar: dot(dA, E)
            const float ds = -(rinv * rinv) * s;                               // VJP of recip, broadcast by sumRows' VJP
            for (int64_t c = lane; c < cols; c += 32) {
                const float e = __expf(z[c]);
                float d;
                if (MODE == 2) { const float a = e * rinv; d = -(aux[r * cols + c] / a); } else d = aux[r * cols + c];
                const float dz = (ds + d * rinv) * e;                          // duplicate's sumT [d1,d2], then map exp VJP
                dZ[r * cols + c] = dz;
                if (MODE == 2) col_acc += dz;
            }
        }
    }
    if (MODE == 2) {                                   // block-level sums first: one red per block for the loss and per column for db
        __shared__ float sh[kThreads / 32][33];        // (same-address reds serialise in L2: one per warp cost ~20 us at 32768 rows)
        loss_acc = warp_sum(loss_acc);
        sh[threadIdx.x >> 5][lane] = col_acc;
        if (lane == 0) sh[threadIdx.x >> 5][32] = loss_acc;
        __syncthreads();
        if (threadIdx.x < 32) {
            float s = 0.f, l = 0.f;
#pragma unroll
            for (int w = 0; w < kThreads / 32; ++w) { s += sh[w][threadIdx.x]; l += sh[w][32]; }
            if (colsum != nullptr && threadIdx.x < cols) atomicAdd(colsum + threadIdx.x, s);
            if (loss != nullptr && threadIdx.x == 0) atomicAdd(loss, l);
        }
    }
}
// softmax + crossEntropy head for few classes (cols <= CMAX): one THREAD per row — a warp per row leaves 22 of 32 lanes idle at
// 10 classes and pays two shuffle reductions per row.  Same arithmetic as k_softmax_rows<2>; db = column sums of dZ.
template <int CMAX>
__global__ void k_softmax_ce_small(const float* __restrict__ Z, const float* __restrict__ Y, float* __restrict__ A, float* __restrict__ dZ,
                                   float* __restrict__ loss, int64_t rows, int cols, float* __restrict__ colsum) {
    const int lane = threadIdx.x & 31;
    float loss_acc = 0.f, col_acc[CMAX];
#pragma unroll
    for (int c = 0; c < CMAX; ++c) col_acc[c] = 0.f;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float e[CMAX], y[CMAX];
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            const bool ok = c < cols;
            e[c] = ok ? __expf(Z[r * cols + c]) : 0.f;
            y[c] = ok ? Y[r * cols + c] : 0.f;
            se += e[c];
        }
        const float rinv = 1.0f / se;
        float s = 0.f, d[CMAX];
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < cols) {
                const float a = e[c] * rinv;
                A[r * cols + c] = a;
                loss_acc -= __logf(a) * y[c];
                d[c] = -(y[c] / a);
                s += d[c] * e[c];
            } else d[c] = 0.f;
        }
        const float ds = -(rinv * rinv) * s;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            if (c < cols) {
                const float dz = (ds + d[c] * rinv) * e[c];
                dZ[r * cols + c] = dz;
                col_acc[c] += dz;
            }
        }
    }
    __shared__ float sh[kThreads / 32][CMAX + 1];
    loss_acc = warp_sum(loss_acc);
#pragma unroll
    for (int c = 0; c < CMAX; ++c) {
        const float v = warp_sum(col_acc[c]);
        if (lane == 0) sh[threadIdx.x >> 5][c] = v;
    }
    if (lane == 0) sh[threadIdx.x >> 5][CMAX] = loss_acc;
    __syncthreads();
    if (threadIdx.x <= CMAX) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += sh[w][threadIdx.x];
        if (threadIdx.x == CMAX) { if (loss != nullptr) atomicAdd(loss, t); }
        else if (colsum != nullptr && threadIdx.x < cols) atomicAdd(colsum + threadIdx.x, t);
    }
}
// softmax map / its VJP on rows of <= CMAX columns: one thread per row (see k_softmax_ce_small), 16-byte accesses when cols % 4 == 0
template <int CMAX, int MODE>   // MODE 0: out = softmax(Z);  MODE 1: out = VJP of softmax at Z applied to dA
__global__ void k_softmax_small(const float* __restrict__ Z, const float* __restrict__ dA, float* __restrict__ out, int64_t rows, int cols, int vec) {
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
        float e[CMAX], d[CMAX];
        const float* z = Z + r * cols;
        const float* da = MODE == 1 ? dA + r * cols : nullptr;
        if (vec) {
#pragma unroll
            for (int c4 = 0; c4 < CMAX / 4; ++c4) {
                if (c4 * 4 < cols) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(z) + c4);
                    e[c4 * 4] = t.x; e[c4 * 4 + 1] = t.y; e[c4 * 4 + 2] = t.z; e[c4 * 4 + 3] = t.w;
                    if (MODE == 1) { const float4 u = __ldg(reinterpret_cast<const float4*>(da) + c4); d[c4 * 4] = u.x; d[c4 * 4 + 1] = u.y; d[c4 * 4 + 2] = u.z; d[c4 * 4 + 3] = u.w; }
                }
            }
        } else {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) if (c < cols) { e[c] = __ldg(z + c); if (MODE == 1) d[c] = __ldg(da + c); }
        }
        float se = 0.f, s = 0.f;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) {
            e[c] = c < cols ? __expf(e[c]) : 0.f;          // map exp (no max-subtraction: NeuralNet.hs:52-59)
            se += e[c];
            if (MODE == 1) s += c < cols ? d[c] * e[c] : 0.f;
        }
        const float rinv = 1.0f / se, ds = -(rinv * rinv) * s;
#pragma unroll
        for (int c = 0; c < CMAX; ++c) e[c] = MODE == 0 ? e[c] * rinv : (c < cols ? (ds + d[c] * rinv) * e[c] : 0.f);
        float* o = out + r * cols;
        if (vec) {
#pragma unroll
            for (int c4 = 0; c4 < CMAX / 4; ++c4)
                if (c4 * 4 < cols) reinterpret_cast<float4*>(o)[c4] = make_float4(e[c4 * 4], e[c4 * 4 + 1], e[c4 * 4 + 2], e[c4 * 4 + 3]);
        } else {
#pragma unroll
            for (int c = 0; c < CMAX; ++c) if (c < cols) o[c] = e[c];
        }
    }
}
__global__ void k_loss_vjp(int loss, const float* __restrict__ A, const float* __restrict__ Y, float* __restrict__ dA, float* __restrict__ out, int64_t n) {
    __shared__ float sh[32];
    float acc = 0.f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float a = A[i], y = Y[i];
        if (loss == 1) { const float d = y - a; acc = fmaf(d, d, acc); dA[i] = -2.0f * d; }
        else { acc -= __logf(a) * y; dA[i] = -(y / a); }
    }
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0 && out) atomicAdd(out, acc);
}

// ------------------------------------------------------------------ CUDA-core GEMM (64x64x16 tiles, 4x4 per thread)
__device__ __forceinline__ void epi_elem(const GemmParams& p, int row, int col, float acc, float& loss_acc) {
    float* o0 = reinterpret_cast<float*>(p.out0) + (long long)row * p.ld_out0 + col;
    switch (p.epi) {
        case EPI_STORE: {
            float v = p.alpha * acc;
            if (p.aux0) v = fmaf(p.beta, reinterpret_cast<const float*>(p.aux0)[(long long)row * p.ld_aux0 + col], v);
            *o0 = v;
        } break;
        case EPI_ATOMIC: atomicAdd(o0, p.alpha * acc); break;
        case EPI_BIAS_ACT:
        case EPI_BIAS_ACT_DZ:
        case EPI_BIAS_ACT_SE: {
            const float a = act_apply(p.act, acc + (p.bias ? p.bias[col] : 0.f));
            *o0 = a;
            if (p.epi != EPI_BIAS_ACT) {
                const float x = reinterpret_cast<const float*>(p.aux0)[(long long)row * p.ld_aux0 + col];
                float* o1 = reinterpret_cast<float*>(p.out1) + (long long)row * p.ld_out1 + col;
                if (p.epi == EPI_BIAS_ACT_DZ) *o1 = x * act_deriv_from_out(p.act, a);
                else { const float d = x - a; loss_acc = fmaf(d, d, loss_acc); *o1 = -2.0f * d * act_deriv_from_out(p.act, a); }
            }
        } break;
        case EPI_MUL_DACT: {
            const float x = reinterpret_cast<const float*>(p.aux0)[(long long)row * p.ld_aux0 + col];
            *o0 = acc * act_deriv_from_out(p.act, x);
        } break;
        default: break;
    }
}

template <int MA, int MB>
__global__ void __launch_bounds__(256) k_gemm_simt(const float* __restrict__ A, long long lda, const float* __restrict__ B, long long ldb, GemmParams p) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    const int kchunk = (p.K + gridDim.z - 1) / gridDim.z;
    const int k_begin = blockIdx.z * kchunk, k_end = min(p.K, k_begin + kchunk);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[4][4] = {};
    for (int k0 = k_begin; k0 < k_end; k0 += 16) {
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            int mm, kk;
            if (MA == MAJOR_K) { kk = i & 15; mm = i >> 4; } else { mm = i & 63; kk = i >> 6; }
            const int gm = m0 + mm, gk = k0 + kk;
            float v = 0.f;
            if (gm < p.M && gk < k_end) v = MA == MAJOR_K ? A[(long long)gm * lda + gk] : A[(long long)gk * lda + gm];
            As[kk][mm] = v;
        }
        for (int i = threadIdx.x; i < 64 * 16; i += 256) {
            int nn, kk;
            if (MB == MAJOR_K) { kk = i & 15; nn = i >> 4; } else { nn = i & 63; kk = i >> 6; }
            const int gn = n0 + nn, gk = k0 + kk;
            float v = 0.f;
            if (gn < p.N && gk < k_end) v = MB == MAJOR_K ? B[(long long)gn * ldb + gk] : B[(long long)gk * ldb + gn];
            Bs[kk][nn] = v;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; b[i] = Bs[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    float loss_acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int row = m0 + ty * 4 + i, col = n0 + tx * 4 + j;
            if (row < p.M && col < p.N) epi_elem(p, row, col, acc[i][j], loss_acc);
        }
    if (p.epi == EPI_BIAS_ACT_SE && p.loss) {
        __shared__ float sh[32];
        loss_acc = block_sum(loss_acc, sh);
        if (threadIdx.x == 0) atomicAdd(p.loss, loss_acc);
    }
}

}  // namespace

// ---------------------------------------------------------------------- contraction followed by sumRows, fused
// out[r,n] = sum_a sum_k x[a,r,k] y[k,n]  ==  (sum_a x[a,r,:]) . y      — `gmul lM 1 lN >>> sumRows` (TOp.hs:56-94,151-159) as ONE pass
// over x: the row sum commutes with the contraction, so the [A,R,N] intermediate of the reference (64 separate 64^3 gemms at
// BASELINE configs[4]) is never formed.  One CTA per r; K*(N+1) floats of y staged in shared memory by the VJP.
constexpr int kGsrThreads = 512;

__global__ void __launch_bounds__(kGsrThreads) k_gsr_fwd(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                                         int64_t A, int64_t R, int K, int N) {
    extern __shared__ float sh[];                 // [groups][K] partial row sums, then xs[K] in sh[0..K)
    const int r = blockIdx.x, t = threadIdx.x;
    const int KT = K < kGsrThreads ? K : kGsrThreads;      // threads along k
    const int groups = kGsrThreads / KT;                    // thread groups along a
    const int g = t / KT, kk = t % KT;
    for (int k0 = 0; k0 < K; k0 += KT) {
        const int k = k0 + kk;
        float acc = 0.f;
        if (g < groups && k < K) {
            const float* p = x + ((int64_t)g * R + r) * K + k;
            const int64_t step = (int64_t)groups * R * K;
            int64_t a = g;
            for (; a + 7 * groups < A; a += 8 * groups, p += 8 * step) {      // 8 independent loads in flight per thread
                const float v0 = __ldg(p), v1 = __ldg(p + step), v2 = __ldg(p + 2 * step), v3 = __ldg(p + 3 * step);
                const float v4 = __ldg(p + 4 * step), v5 = __ldg(p + 5 * step), v6 = __ldg(p + 6 * step), v7 = __ldg(p + 7 * step);
                acc += ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
            }
            for (; a < A; a += groups, p += step) acc += __ldg(p);
            sh[g * K + k] = acc;
        }
    }
    __syncthreads();
    for (int k = t; k < K; k += kGsrThreads) {
        float s = sh[k];
        for (int gg = 1; gg < groups; ++gg) s += sh[gg * K + k];
        sh[k] = s;
    }
    __syncthreads();
    for (int n = t; n < N; n += kGsrThreads) {
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc = fmaf(sh[k], __ldg(y + (int64_t)k * N + n), acc);
        out[(int64_t)r * N + n] = acc;
    }
}

// The same with y staged in shared memory (K*N floats) by loads that are in flight together with the loads of x — one memory latency
// instead of one per k of the final contraction — and the contraction spread over all threads (N columns x threads/N slices of k).
constexpr int kGsrYB = 8;   // y elements per thread whose loads are issued before the x loads (config 5: all of them)
__global__ void __launch_bounds__(kGsrThreads) k_gsr_fwd_staged(const float* __restrict__ x, const float* __restrict__ y, float* __restrict__ out,
                                                                int64_t A, int64_t R, int K, int N) {
    extern __shared__ float sh[];
    const int KT = K < kGsrThreads ? K : kGsrThreads;
    const int groups = kGsrThreads / KT;
    const int G3 = N < kGsrThreads ? kGsrThreads / N : 1;   // k slices of the contraction
    float* ys = sh;                                          // [K][N]
    float* xp = ys + (size_t)K * N;                          // [groups][K] -> xs[K]
    float* red = xp + (size_t)groups * K;                    // [G3][N]
    const int r = blockIdx.x, t = threadIdx.x;
    const int g = t / KT, kk = t % KT;
    float yv[kGsrYB];
#pragma unroll
    for (int j = 0; j < kGsrYB; ++j) { const int i = t + j * kGsrThreads; yv[j] = i < K * N ? __ldg(y + i) : 0.f; }
    for (int k0 = 0; k0 < K; k0 += KT) {
        const int k = k0 + kk;
        float acc = 0.f;
        if (g < groups && k < K) {
            const float* p = x + ((int64_t)g * R + r) * K + k;
            const int64_t step = (int64_t)groups * R * K;
            int64_t a = g;
            for (; a + 7 * groups < A; a += 8 * groups, p += 8 * step) {
                const float v0 = __ldg(p), v1 = __ldg(p + step), v2 = __ldg(p + 2 * step), v3 = __ldg(p + 3 * step);
                const float v4 = __ldg(p + 4 * step), v5 = __ldg(p + 5 * step), v6 = __ldg(p + 6 * step), v7 = __ldg(p + 7 * step);
                acc += ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
            }
            for (; a < A; a += groups, p += step) acc += __ldg(p);
            xp[g * K + k] = acc;
        }
    }
#pragma unroll
    for (int j = 0; j < kGsrYB; ++j) { const int i = t + j * kGsrThreads; if (i < K * N) ys[i] = yv[j]; }
    for (int i = t + kGsrYB * kGsrThreads; i < K * N; i += kGsrThreads) ys[i] = __ldg(y + i);
    __syncthreads();
    for (int k = t; k < K; k += kGsrThreads) {
        float s_ = xp[k];
        for (int gg = 1; gg < groups; ++gg) s_ += xp[gg * K + k];
        xp[k] = s_;
    }
    __syncthreads();
    if (N < kGsrThreads) {
        const int n = t % N, kg = t / N;
        if (kg < G3) {
            float acc = 0.f;
            for (int k = kg; k < K; k += G3) acc = fmaf(xp[k], ys[k * N + n], acc);
            red[kg * N + n] = acc;
        }
        __syncthreads();
        if (t < N) {
            float acc = red[t];
            for (int q = 1; q < G3; ++q) acc += red[q * N + t];
            out[(int64_t)r * N + t] = acc;
        }
    } else {
        for (int n = t; n < N; n += kGsrThreads) {
            float acc = 0.f;
            for (int k = 0; k < K; ++k) acc = fmaf(xp[k], ys[k * N + n], acc);
            out[(int64_t)r * N + n] = acc;
        }
    }
}
size_t gsr_fwd_staged_smem(int K, int N) {
    const int KT = K < kGsrThreads ? K : kGsrThreads;
    return sizeof(float) * ((size_t)K * N + (size_t)(kGsrThreads / KT) * K + (size_t)(N < kGsrThreads ? kGsrThreads / N : 1) * N);
}

// VJP with cotangent ct[R,N]:  dx[a,r,k] = sum_n ct[r,n] y[k,n]  (the same for every a: written A times, never materialising the
// broadcast cotangent),  dy[k,n] += (sum_a x[a,r,k]) ct[r,n]  (rank-1 update per r, fp32 reds into the pre-zeroed dy).
template <int PER>   // PER > 0: dy partials of RB rows kept in PER registers per thread (K*N <= PER * threads); 0: one red per row and entry
__global__ void __launch_bounds__(kGsrThreads) k_gsr_vjp(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ ct,
                                                         float* __restrict__ dx, float* __restrict__ dy, int64_t A, int64_t R, int K, int N, int RB) {
    extern __shared__ float sh[];
    float* ys = sh;                               // [K][N+1] (padded: a thread per k walks n)
    float* cts = ys + (size_t)K * (N + 1);        // [N]
    float* tk = cts + N;                          // [K]   t[k] = sum_n ct[r,n] y[k,n]
    float* xs = tk + K;                           // [groups][K] -> xs[K]
    const int t = threadIdx.x;
    float yv[kGsrYB];                             // the first kGsrYB * threads elements of y: in flight together with the first row's x loads
#pragma unroll
    for (int j = 0; j < kGsrYB; ++j) { const int i = t + j * kGsrThreads; yv[j] = i < K * N ? __ldg(y + i) : 0.f; }
    bool y_staged = false;
    const int KT = K < kGsrThreads ? K : kGsrThreads;
    const int groups = kGsrThreads / KT;
    const int g = t / KT, kk = t % KT;
    float part[PER > 0 ? PER : 1];
#pragma unroll
    for (int j = 0; j < (PER > 0 ? PER : 1); ++j) part[j] = 0.f;
    const int64_t r_end = min(R, (int64_t)(blockIdx.x + 1) * RB);
    for (int64_t r = (int64_t)blockIdx.x * RB; r < r_end; ++r) {
        __syncthreads();                          // previous row's xs / cts / tk fully consumed
        for (int n = t; n < N; n += kGsrThreads) cts[n] = __ldg(ct + r * N + n);
        for (int k0 = 0; k0 < K; k0 += KT) {      // row sums of x over a (needed by dy), coalesced along k
            const int k = k0 + kk;
            float acc = 0.f;
            if (g < groups && k < K) {
                const float* p = x + ((int64_t)g * R + r) * K + k;
                const int64_t step = (int64_t)groups * R * K;
                int64_t a = g;
                for (; a + 7 * groups < A; a += 8 * groups, p += 8 * step) {
                    const float v0 = __ldg(p), v1 = __ldg(p + step), v2 = __ldg(p + 2 * step), v3 = __ldg(p + 3 * step);
                    const float v4 = __ldg(p + 4 * step), v5 = __ldg(p + 5 * step), v6 = __ldg(p + 6 * step), v7 = __ldg(p + 7 * step);
                    acc += ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
                }
                for (; a < A; a += groups, p += step) acc += __ldg(p);
                xs[g * K + k] = acc;
            }
        }
        if (!y_staged) {
#pragma unroll
            for (int j = 0; j < kGsrYB; ++j) { const int i = t + j * kGsrThreads; if (i < K * N) ys[(i / N) * (N + 1) + (i % N)] = yv[j]; }
            for (int i = t + kGsrYB * kGsrThreads; i < K * N; i += kGsrThreads) ys[(i / N) * (N + 1) + (i % N)] = __ldg(y + i);
            y_staged = true;
        }
        __syncthreads();
        for (int k = t; k < K; k += kGsrThreads) {
            float s = xs[k];
            for (int gg = 1; gg < groups; ++gg) s += xs[gg * K + k];
            xs[k] = s;
            float acc = 0.f;
            for (int n = 0; n < N; ++n) acc = fmaf(cts[n], ys[k * (N + 1) + n], acc);
            tk[k] = acc;
        }
        __syncthreads();
        for (int k0 = 0; k0 < K; k0 += KT) {      // dx: A copies of t[k], coalesced along k
            const int k = k0 + kk;
            if (g < groups && k < K) {
                const float v = tk[k];
                for (int64_t a = g; a < A; a += groups) dx[(a * R + r) * K + k] = v;
            }
        }
        if constexpr (PER > 0) {
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int i = t + j * kGsrThreads;
                if (i < K * N) part[j] = fmaf(xs[i / N], cts[i % N], part[j]);
            }
        } else {
            for (int i = t; i < K * N; i += kGsrThreads) atomicAdd(dy + i, xs[i / N] * cts[i % N]);
        }
    }
    if constexpr (PER > 0) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int i = t + j * kGsrThreads;
            if (i < K * N) atomicAdd(dy + i, part[j]);
        }
    }
}

// The same VJP without reds and without a zeroed dy: block b does two independent jobs —
//   (b < R)  dx[:, b, :] = t[k] = sum_n ct[b,n] y[k,n]              (as above)
//   (b < K)  dy[b, :]    = sum_r (sum_a x[a,r,b]) ct[r,:]            (column b of the row sums, read with stride K: 32-byte sectors from
//                                                                   L2, where the forward kernel has just left x)
// so every element of dy is written once, deterministically, and the memset node and the tail of 64-way contended reds disappear
// (config 5 as a recorded step: 12.3 -> see profiles).  All global loads of a block are issued before its first barrier.
__global__ void __launch_bounds__(kGsrThreads) k_gsr_vjp_det(const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ ct,
                                                             float* __restrict__ dx, float* __restrict__ dy, int64_t A, int64_t R, int K, int N) {
    extern __shared__ float sh[];
    const int G = N < kGsrThreads ? kGsrThreads / N : 1;   // slices (of r, of n) that share a column in the two contractions
    const int RT = (int)(R < kGsrThreads ? R : kGsrThreads), GA = kGsrThreads / RT;    // job 2: RT threads along r, GA groups along a
    float* ys = sh;                                   // [K][N+1]
    float* cts = ys + (size_t)K * (N + 1);            // [N]      ct[b,:]
    float* tk = cts + N;                              // [K]
    float* ctall = tk + K;                            // [R][N]
    float* xpart = ctall + (size_t)R * N;             // [GA][R] -> xs[R]
    float* red = xpart + (size_t)GA * R;              // [G][N]
    const int b = blockIdx.x, t = threadIdx.x;
    const bool job1 = b < R, job2 = b < K;
    // ---- loads (job 2's strided column of x first: the longest chain)
    if (job2) {
        const int rr = t % RT, ga = t / RT;
        for (int64_t r0 = 0; r0 < R; r0 += RT) {
            const int64_t r = r0 + rr;
            float acc = 0.f;
            if (ga < GA && r < R) {
                const float* p = x + ((int64_t)ga * R + r) * K + b;
                const int64_t step = (int64_t)GA * R * K;
                int64_t a = ga;
                for (; a + 7 * GA < A; a += 8 * GA, p += 8 * step) {
                    const float v0 = __ldg(p), v1 = __ldg(p + step), v2 = __ldg(p + 2 * step), v3 = __ldg(p + 3 * step);
                    const float v4 = __ldg(p + 4 * step), v5 = __ldg(p + 5 * step), v6 = __ldg(p + 6 * step), v7 = __ldg(p + 7 * step);
                    acc += ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
                }
                for (; a < A; a += GA, p += step) acc += __ldg(p);
                xpart[ga * R + r] = acc;
            }
        }
        for (int64_t i = t; i < R * N; i += kGsrThreads) ctall[i] = __ldg(ct + i);
    }
    if (job1) {
        for (int i = t; i < K * N; i += kGsrThreads) ys[(i / N) * (N + 1) + (i % N)] = __ldg(y + i);
        for (int n = t; n < N; n += kGsrThreads) cts[n] = __ldg(ct + (int64_t)b * N + n);
    }
    __syncthreads();
    // ---- job 1: t[k], then dx
    if (job1) {
        for (int k = t; k < K; k += kGsrThreads) {
            float acc = 0.f;
            for (int n = 0; n < N; ++n) acc = fmaf(cts[n], ys[k * (N + 1) + n], acc);
            tk[k] = acc;
        }
    }
    // ---- job 2: xs[r] = sum over the a-groups
    if (job2) {
        for (int r = t; r < R; r += kGsrThreads) {
            float s_ = xpart[r];
            for (int g = 1; g < GA; ++g) s_ += xpart[g * R + r];
            xpart[r] = s_;
        }
    }
    __syncthreads();
    if (job1) {
        const int KT = K < kGsrThreads ? K : kGsrThreads, groups = kGsrThreads / KT;
        const int g = t / KT, kk = t % KT;
        for (int k0 = 0; k0 < K; k0 += KT) {
            const int k = k0 + kk;
            if (g < groups && k < K) {
                const float v = tk[k];
                for (int64_t a = g; a < A; a += groups) dx[(a * R + b) * K + k] = v;
            }
        }
    }
    if (job2) {   // uniform per block: the barrier below is reached by all threads or none
        if (N < kGsrThreads) {
            const int n = t % N, rg = t / N;
            if (rg < G) {
                float acc = 0.f;
                for (int64_t r = rg; r < R; r += G) acc = fmaf(xpart[r], ctall[r * N + n], acc);
                red[rg * N + n] = acc;
            }
            __syncthreads();
            if (t < N) {
                float acc = red[t];
                for (int q = 1; q < G; ++q) acc += red[q * N + t];
                dy[(int64_t)b * N + t] = acc;
            }
        } else {
            for (int n = t; n < N; n += kGsrThreads) {
                float acc = 0.f;
                for (int64_t r = 0; r < R; ++r) acc = fmaf(xpart[r], ctall[r * N + n], acc);
                dy[(int64_t)b * N + n] = acc;
            }
        }
    }
}
size_t gsr_vjp_det_smem(int64_t R, int K, int N) {
    const int64_t RT = R < kGsrThreads ? R : kGsrThreads;
    return sizeof(float) * ((size_t)K * (N + 1) + N + K + (size_t)R * N + (size_t)(kGsrThreads / RT) * R + (size_t)(N < kGsrThreads ? kGsrThreads / N : 1) * N);
}

// ====================================================================== launch wrappers
namespace {
__global__ void k_mc_push(const float* __restrict__ src, float* out_mc, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) ptx::multimem_red_add(out_mc + i, src[i]);
}
}  // namespace
void mc_push(const LaunchCtx& lc, const float* src, float* out_mc, int64_t n) {
    if (n <= 0) return;
    k_mc_push<<<grid_for(lc, n), kThreads, 0, lc.stream>>>(src, out_mc, n);
    count(lc);
}

namespace {
__global__ void k_split_bf16(const float* __restrict__ x, __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ lo, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float a = x[i];
        v[i] = __float2bfloat16_rn(a);
        lo[i] = __float2bfloat16_rn(a - __uint_as_float(__float_as_uint(a) & 0xffffe000u));
    }
}
}  // namespace
void split_bf16(const LaunchCtx& lc, const float* x, void* v, void* lo, int64_t n) {
    if (n <= 0) return;
    k_split_bf16<<<grid_for(lc, n), kThreads, 0, lc.stream>>>(x, (__nv_bfloat16*)v, (__nv_bfloat16*)lo, n);
    count(lc);
}

size_t gsr_vjp_smem(int K, int N) { return sizeof(float) * ((size_t)K * (N + 1) + N + K + (size_t)(kGsrThreads / (K < kGsrThreads ? K : kGsrThreads)) * K); }
bool gsr_fits(int64_t K, int64_t N) { return K >= 1 && N >= 1 && K <= 1024 && N <= 4096 && gsr_vjp_smem((int)K, (int)N) <= 96 * 1024; }
void gsr_fwd(const LaunchCtx& lc, const float* x, const float* y, float* out, int64_t A, int64_t R, int K, int N) {
    const int KT = K < kGsrThreads ? K : kGsrThreads;
    const size_t staged = gsr_fwd_staged_smem(K, N);
    if (staged <= 96 * 1024) {
        static unsigned long long done_mask = 0;      // opt-in above 48 KiB, once per device
        int dev = 0;
        cudaGetDevice(&dev);
        if (staged > 48 * 1024 && (dev >= 64 || !((done_mask >> dev) & 1ull))) {
            cudaFuncSetAttribute(k_gsr_fwd_staged, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (dev < 64) done_mask |= 1ull << dev;
        }
        k_gsr_fwd_staged<<<(unsigned)R, kGsrThreads, staged, lc.stream>>>(x, y, out, A, R, K, N);
        count(lc);
        return;
    }
    k_gsr_fwd<<<(unsigned)R, kGsrThreads, sizeof(float) * (size_t)(kGsrThreads / KT) * K, lc.stream>>>(x, y, out, A, R, K, N);
    count(lc);
}
bool gsr_vjp_needs_zeroed_dy(int64_t R, int K, int N) { return !(R <= 4096 && gsr_vjp_det_smem(R, K, N) <= 96 * 1024); }
void gsr_vjp(const LaunchCtx& lc, const float* x, const float* y, const float* ct, float* dx, float* dy, int64_t A, int64_t R, int K, int N) {
    if (!gsr_vjp_needs_zeroed_dy(R, K, N)) {
        const size_t smem_det = gsr_vjp_det_smem(R, K, N);
        static unsigned long long det_mask = 0;       // opt-in above 48 KiB, once per device
        int dev = 0;
        cudaGetDevice(&dev);
        if (smem_det > 48 * 1024 && (dev >= 64 || !((det_mask >> dev) & 1ull))) {
            cudaFuncSetAttribute(k_gsr_vjp_det, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
            if (dev < 64) det_mask |= 1ull << dev;
        }
        const unsigned grid = (unsigned)(R > K ? R : K);
        k_gsr_vjp_det<<<grid, kGsrThreads, smem_det, lc.stream>>>(x, y, ct, dx, dy, A, R, K, N);
        count(lc);
        return;
    }
    const size_t smem = gsr_vjp_smem(K, N);
    static unsigned long long attr_done_mask = 0;     // per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((attr_done_mask >> dev) & 1ull)) {
        cudaFuncSetAttribute(k_gsr_vjp<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        cudaFuncSetAttribute(k_gsr_vjp<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
        if (dev < 64) attr_done_mask |= 1ull << dev;
    }
    // rows per CTA: one, until there are more rows than 4 waves of CTAs (measured at R = 64: 4 rows per CTA is 1.8x SLOWER — the rows
    // of one CTA are serialised latency chains, and the K*N reds per CTA were never the bottleneck)
    int RB = (int)((R + 4 * (int64_t)lc.num_sms - 1) / (4 * (int64_t)lc.num_sms));
    if (RB < 1) RB = 1;
    const unsigned grid = (unsigned)((R + RB - 1) / RB);
    if ((int64_t)K * N <= 8 * kGsrThreads) k_gsr_vjp<8><<<grid, kGsrThreads, smem, lc.stream>>>(x, y, ct, dx, dy, A, R, K, N, RB);
    else k_gsr_vjp<0><<<grid, kGsrThreads, smem, lc.stream>>>(x, y, ct, dx, dy, A, R, K, N, RB);
    count(lc);
}

void fill(const LaunchCtx& lc, float* p, int64_t n, float v) { if (n <= 0) return; k_fill<<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(p, n, v, aligned16(p)); count(lc); }
void fill_bf16(const LaunchCtx& lc, void* p, int64_t n, float v) { if (n <= 0) return; k_fill_bf16<<<grid_for(lc, n), kThreads, 0, lc.stream>>>((__nv_bfloat16*)p, n, v); count(lc); }
void rand_normal(const LaunchCtx& lc, float* p, int64_t n, float mean, float sd, uint64_t seed) { if (n <= 0) return; k_rand<true><<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(p, n, mean, sd, seed); count(lc); }
void rand_uniform(const LaunchCtx& lc, float* p, int64_t n, float lo, float hi, uint64_t seed) { if (n <= 0) return; k_rand<false><<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(p, n, lo, hi, seed); count(lc); }
void cast_f32_bf16(const LaunchCtx& lc, const float* s, void* d, int64_t n) { if (n <= 0) return; k_cast_f2b<<<grid_for(lc, (n + 7) / 8), kThreads, 0, lc.stream>>>(s, (__nv_bfloat16*)d, n, aligned16(s) && aligned16(d)); count(lc); }
void cast_bf16_f32(const LaunchCtx& lc, const void* s, float* d, int64_t n) { if (n <= 0) return; k_cast_b2f<<<grid_for(lc, (n + 7) / 8), kThreads, 0, lc.stream>>>((const __nv_bfloat16*)s, d, n, aligned16(s) && aligned16(d)); count(lc); }
void eye(const LaunchCtx& lc, float* p, int64_t n) { if (n <= 0) return; k_eye<<<grid_for(lc, n * n), kThreads, 0, lc.stream>>>(p, n); count(lc); }

void axpy(const LaunchCtx& lc, float alpha, const float* x, const float* y, float* out, int64_t n) {
    if (n <= 0) return;
    const int g = grid_for(lc, (n + 3) / 4);
    if (y) {
        const bool vec = aligned16(x) && aligned16(y) && aligned16(out);
        k_map2<<<g, kThreads, 0, lc.stream>>>(x, y, out, n, vec, [alpha] __device__(float a, float b) { return fmaf(alpha, a, b); });
    } else {
        const bool vec = aligned16(x) && aligned16(out);
        k_map1<<<g, kThreads, 0, lc.stream>>>(x, out, n, vec, [alpha] __device__(float a) { return alpha * a; });
    }
    count(lc);
}
void add_n(const LaunchCtx& lc, int n_in, const float* const* xs, float* out, int64_t n) {
    if (n <= 0) return;
    InPtrs in{};
    bool vec = aligned16(out);
    for (int j = 0; j < n_in; ++j) { in.p[j] = xs[j]; vec = vec && aligned16(xs[j]); }
    k_add_n<<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(in, n_in, out, n, vec);
    count(lc);
}
void sgd(const LaunchCtx& lc, const float* p, const float* g, float rate, float* out, int64_t n) {
    if (n <= 0) return;
    const bool vec = aligned16(p) && aligned16(g) && aligned16(out);
    k_map2<<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(p, g, out, n, vec, [rate] __device__(float a, float b) { return a - rate * b; });   // FeedForward.hs:141-147
    count(lc);
}
void dact_mul(const LaunchCtx& lc, int act, const float* dA, const float* A, float* dZ, int64_t n) {
    if (n <= 0) return;
    const bool vec = aligned16(dA) && aligned16(A) && aligned16(dZ);
    k_map2<<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(dA, A, dZ, n, vec, [act] __device__(float d, float a) { return d * act_deriv_from_out(act, a); });
    count(lc);
}
void bias_act(const LaunchCtx& lc, int act, const float* Z, const float* bias, float* A, int64_t rows, int64_t cols) {
    if (rows * cols <= 0) return;
    const bool vec = (cols % 4) == 0 && aligned16(Z) && aligned16(A) && (!bias || aligned16(bias));
    const dim3 g = grid_rows_cols(lc, rows, vec ? cols / 4 : cols);
    if (vec) k_bias_act<true><<<g, kThreads, 0, lc.stream>>>(act, Z, bias, A, rows, cols);
    else k_bias_act<false><<<g, kThreads, 0, lc.stream>>>(act, Z, bias, A, rows, cols);
    count(lc);
}
// Catalogue of lifted programs that get a specialised float4 kernel instead of the stack interpreter (HBM-bound instead of
// ALU-bound): exactly the closures the reference's own TOps lift (NeuralNet.hs:38-50 logistic / d * logistic'(x) as `map'` builds
// it, exp / log / recip of softmax and crossEntropy NeuralNet.hs:52-77, the SGD rule p - r*g FeedForward.hs:141-147, the squared
// difference of squaredError) plus every single unary / binary opcode.  Matching is exact on the bytecode.
namespace {
constexpr int32_t ins(int op, int arg = 0) { return (op << 16) | arg; }
bool prog_is(const LiftProgram& p, std::initializer_list<int32_t> code) {
    if (p.len != (int)code.size()) return false;
    int i = 0;
    for (int32_t c : code) if (p.code[i++] != c) return false;
    return true;
}
template <typename F> void launch_map1(const LaunchCtx& lc, const float* x, float* out, int64_t n, F f) {
    k_map1<<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(x, out, n, aligned16(x) && aligned16(out), f); count(lc);
}
template <typename F> void launch_map2(const LaunchCtx& lc, const float* x, const float* y, float* out, int64_t n, F f) {
    k_map2<<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(x, y, out, n, aligned16(x) && aligned16(y) && aligned16(out), f); count(lc);
}
bool lift_catalogue(const LaunchCtx& lc, const LiftProgram& p, int n_in, const float* const* in, float* out, int64_t n) {
    if (p.len == 2 && (p.code[0] >> 16) == 0 && (p.code[0] & 0xffff) < n_in) {          // [VAR a, unary]
        const float* x = in[p.code[0] & 0xffff];
        switch (p.code[1] >> 16) {
            case 6: launch_map1(lc, x, out, n, [] __device__(float a) { return -a; }); return true;
            case 7: launch_map1(lc, x, out, n, [] __device__(float a) { return __expf(a); }); return true;
            case 8: launch_map1(lc, x, out, n, [] __device__(float a) { return __logf(a); }); return true;
            case 9: launch_map1(lc, x, out, n, [] __device__(float a) { return __fdividef(1.0f, a); }); return true;
            case 10: launch_map1(lc, x, out, n, [] __device__(float a) { return sqrtf(a); }); return true;
            case 11: launch_map1(lc, x, out, n, [] __device__(float a) { return tanhf(a); }); return true;
            case 12: launch_map1(lc, x, out, n, [] __device__(float a) { return fabsf(a); }); return true;
            case 17: launch_map1(lc, x, out, n, [] __device__(float a) { return __fdividef(1.0f, 1.0f + __expf(-a)); }); return true;
            default: return false;
        }
    }
    if (n_in >= 1 && prog_is(p, {ins(1, 0), ins(0, 0), ins(5)}) && p.consts[0] == 1.0f) {  // 1 / x
        launch_map1(lc, in[0], out, n, [] __device__(float a) { return __fdividef(1.0f, a); }); return true;
    }
    if (n_in < 2) return false;
    if (p.len == 3 && (p.code[0] >> 16) == 0 && (p.code[1] >> 16) == 0 && (p.code[0] & 0xffff) < n_in && (p.code[1] & 0xffff) < n_in) {   // [VAR a, VAR b, binary]
        const float *x = in[p.code[0] & 0xffff], *y = in[p.code[1] & 0xffff];
        switch (p.code[2] >> 16) {
            case 2: launch_map2(lc, x, y, out, n, [] __device__(float a, float b) { return a + b; }); return true;
            case 3: launch_map2(lc, x, y, out, n, [] __device__(float a, float b) { return a - b; }); return true;
            case 4: launch_map2(lc, x, y, out, n, [] __device__(float a, float b) { return a * b; }); return true;
            case 5: launch_map2(lc, x, y, out, n, [] __device__(float a, float b) { return a / b; }); return true;
            case 14: launch_map2(lc, x, y, out, n, [] __device__(float a, float b) { return fmaxf(a, b); }); return true;
            case 15: launch_map2(lc, x, y, out, n, [] __device__(float a, float b) { return fminf(a, b); }); return true;
            default: return false;
        }
    }
    if (prog_is(p, {ins(0, 0), ins(1, 0), ins(0, 1), ins(4), ins(3)})) {                  // p - r * g          (trainNetwork's zip)
        const float r = p.consts[0];
        launch_map2(lc, in[0], in[1], out, n, [r] __device__(float a, float g) { return a - r * g; }); return true;
    }
    if (prog_is(p, {ins(0, 0), ins(0, 1), ins(17), ins(1, 0), ins(0, 1), ins(17), ins(3), ins(4), ins(4)}) && p.consts[0] == 1.0f) {
        launch_map2(lc, in[0], in[1], out, n, [] __device__(float d, float x) {                // d * logistic'(x)   (map' logistic logistic')
            const float s = __fdividef(1.0f, 1.0f + __expf(-x));
            return d * (s * (1.0f - s));
        });
        return true;
    }
    if (prog_is(p, {ins(0, 0), ins(0, 1), ins(3), ins(0, 0), ins(0, 1), ins(3), ins(4)})) { // (a - b)^2          (squaredError's zip)
        launch_map2(lc, in[0], in[1], out, n, [] __device__(float a, float b) { const float d = a - b; return d * d; }); return true;
    }
    return false;
}
}  // namespace

int64_t g_lift_catalogue_hits = 0;   // instrumentation (tests): programs served by a specialised kernel
void lift(const LaunchCtx& lc, const LiftProgram& prog, int n_in, const float* const* in, float* out, int64_t n) {
    if (n <= 0) return;
    if (lift_catalogue(lc, prog, n_in, in, out, n)) { ++g_lift_catalogue_hits; return; }
    InPtrs ip{};
    bool vec = aligned16(out);
    for (int j = 0; j < n_in; ++j) { ip.p[j] = in[j]; vec = vec && aligned16(in[j]); }
    k_lift<<<grid_for(lc, (n + 3) / 4), kThreads, 0, lc.stream>>>(prog, ip, n_in, out, n, vec);
    count(lc);
}

void sum_all(const LaunchCtx& lc, const float* x, int64_t n, float* out, float* ws) {
    const int g = grid_for(lc, n > 0 ? (n + 3) / 4 : 1, kThreads, 4);
    k_reduce_partial<false><<<g, kThreads, 0, lc.stream>>>(x, nullptr, n, ws); count(lc);
    k_reduce_final<<<1, kThreads, 0, lc.stream>>>(ws, g, out); count(lc);
}
void dot(const LaunchCtx& lc, const float* x, const float* y, int64_t n, float* out, float* ws) {
    const int g = grid_for(lc, n > 0 ? (n + 3) / 4 : 1, kThreads, 4);
    k_reduce_partial<true><<<g, kThreads, 0, lc.stream>>>(x, y, n, ws); count(lc);
    k_reduce_final<<<1, kThreads, 0, lc.stream>>>(ws, g, out); count(lc);
}
void trace(const LaunchCtx& lc, const float* a, int64_t n, int64_t ld, float* out) { k_trace<<<1, kThreads, 0, lc.stream>>>(a, n, ld, out); count(lc); }

static int colsum_chunks(const LaunchCtx& lc, int64_t rows, int64_t cols) {
    const int64_t cb = (cols + 127) / 128;
    int64_t chunks = (4 * (int64_t)lc.num_sms + cb - 1) / cb;
    if (chunks > 64) chunks = 64;
    if (chunks > (rows + 7) / 8) chunks = (rows + 7) / 8;
    if (chunks < 1) chunks = 1;
    return (int)chunks;
}
void col_sums(const LaunchCtx& lc, const float* x, int64_t rows, int64_t cols, float* out, float* ws, bool accumulate) {
    if (cols <= 0) return;
    const int chunks = colsum_chunks(lc, rows, cols);
    const int64_t rpc = (rows + chunks - 1) / chunks;
    dim3 g((unsigned)((cols + 127) / 128), (unsigned)chunks);
    k_colsum_partial<float><<<g, 256, 0, lc.stream>>>(x, nullptr, rows, cols, rpc, ws); count(lc);
    k_colsum_final<<<grid_for(lc, cols), kThreads, 0, lc.stream>>>(ws, chunks, cols, 1.0f, 1.0f, accumulate ? out : nullptr, out); count(lc);
}
void col_sums_bf16(const LaunchCtx& lc, const void* x, int64_t rows, int64_t cols, float* out, float* ws, bool accumulate) {
    if (cols <= 0) return;
    const int chunks = colsum_chunks(lc, rows, cols);
    const int64_t rpc = (rows + chunks - 1) / chunks;
    dim3 g((unsigned)((cols + 127) / 128), (unsigned)chunks);
    k_colsum_partial<__nv_bfloat16><<<g, 256, 0, lc.stream>>>((const __nv_bfloat16*)x, nullptr, rows, cols, rpc, ws); count(lc);
    k_colsum_final<<<grid_for(lc, cols), kThreads, 0, lc.stream>>>(ws, chunks, cols, 1.0f, 1.0f, accumulate ? out : nullptr, out); count(lc);
}

void ger(const LaunchCtx& lc, const float* x, const float* y, float* out, int64_t n, int64_t m) {
    if (n * m <= 0) return;
    const bool vec = (m % 4) == 0 && aligned16(y) && aligned16(out);
    const dim3 g = grid_rows_cols(lc, n, vec ? m / 4 : m);
    if (vec) k_outer_rows<true><<<g, kThreads, 0, lc.stream>>>(x, y, out, n, m);
    else k_outer_rows<false><<<g, kThreads, 0, lc.stream>>>(x, y, out, n, m);
    count(lc);
}
void gemv(const LaunchCtx& lc, float alpha, const float* a, int a_tr, const float* x, float beta, const float* y, float* out, int64_t n, int64_t m) {
    if (n <= 0) return;
    if (!a_tr) {
        const bool vec = aligned16(a) && aligned16(x) && (m % 4 == 0);
        k_gemv_rows<<<grid_for(lc, n * 32, kThreads, 8), kThreads, 0, lc.stream>>>(alpha, a, x, beta, y, out, n, m, vec);
        count(lc);
    } else {
        // A stored [m, n]: weighted column sums, two-stage; workspace comes from the caller through `out`-sized temp — use a static per-stream scratch
        // (rows = m, cols = n).  chunks*n floats of scratch are carved from a cudaMallocAsync allocation.
        const int chunks = colsum_chunks(lc, m, n);
        const int64_t rpc = (m + chunks - 1) / chunks;
        float* ws = nullptr;
        cudaMallocAsync(&ws, sizeof(float) * (size_t)chunks * (size_t)n, lc.stream);
        dim3 g((unsigned)((n + 127) / 128), (unsigned)chunks);
        k_colsum_partial<float><<<g, 256, 0, lc.stream>>>(a, x, m, n, rpc, ws); count(lc);
        k_colsum_final<<<grid_for(lc, n), kThreads, 0, lc.stream>>>(ws, chunks, n, alpha, beta, y, out); count(lc);
        cudaFreeAsync(ws, lc.stream);
    }
}
void transpose2d(const LaunchCtx& lc, const float* in, float* out, int64_t rows, int64_t cols) {
    if (rows * cols <= 0) return;
    dim3 g((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), b(32, 8);
    k_transpose<<<g, b, 0, lc.stream>>>(in, out, rows, cols); count(lc);
}
void permute(const LaunchCtx& lc, const float* in, float* out, int rank, const int64_t* in_dims, const int* perm) {
    PermArgs pa{};
    pa.rank = rank;
    int64_t in_strides[8];
    int64_t s = 1;
    for (int a = rank - 1; a >= 0; --a) { in_strides[a] = s; s *= in_dims[a]; }
    const int64_t n = s;
    for (int a = 0; a < rank; ++a) { pa.out_dims[a] = in_dims[perm[a]]; pa.in_strides_for_out[a] = in_strides[perm[a]]; }
    if (n <= 0) return;
    // output's fastest axis differs from the input's: tile over (input's last axis p, output's last axis q) through shared memory
    if (rank >= 2 && perm[rank - 1] != rank - 1) {
        PermTiled pt{};
        int64_t out_strides[8];
        int64_t so = 1;
        for (int a = rank - 1; a >= 0; --a) { out_strides[a] = so; so *= pa.out_dims[a]; }
        int p_out_axis = -1;                                   // where the input's last axis lands in the output
        for (int a = 0; a < rank; ++a) if (perm[a] == rank - 1) p_out_axis = a;
        pt.P = in_dims[rank - 1]; pt.Q = pa.out_dims[rank - 1];
        pt.p_out_stride = out_strides[p_out_axis]; pt.q_in_stride = pa.in_strides_for_out[rank - 1];
        int64_t batch = 1;
        for (int a = 0; a < rank - 1; ++a) {                   // every output axis except q and p's landing place
            if (a == p_out_axis) continue;
            pt.bdim[pt.nb] = pa.out_dims[a]; pt.bin[pt.nb] = pa.in_strides_for_out[a]; pt.bout[pt.nb] = out_strides[a]; ++pt.nb;
            batch *= pa.out_dims[a];
        }
        const int64_t gx = (pt.P + 31) / 32, gy = (pt.Q + 31) / 32;
        if (gy <= 65535 && batch <= 65535) {
            k_permute_tiled<<<dim3((unsigned)gx, (unsigned)gy, (unsigned)batch), dim3(32, 8), 0, lc.stream>>>(in, out, pt); count(lc);
            return;
        }
    }
    k_permute<<<grid_for(lc, n), kThreads, 0, lc.stream>>>(in, out, pa, n); count(lc);
}
void broadcast_rows(const LaunchCtx& lc, const float* row, float* out, int64_t n, int64_t m) {
    if (n * m <= 0) return;
    const bool vec = (m % 4) == 0 && aligned16(row) && aligned16(out);
    const dim3 g = grid_rows_cols(lc, n, vec ? m / 4 : m);
    if (vec) k_outer_rows<true><<<g, kThreads, 0, lc.stream>>>(nullptr, row, out, n, m);
    else k_outer_rows<false><<<g, kThreads, 0, lc.stream>>>(nullptr, row, out, n, m);
    count(lc);
}
static int64_t diag_step(int64_t n, int rank) { int64_t step = 0, s = 1; for (int a = 0; a < rank; ++a) { step += s; s *= n; } return step; }
void diag_embed(const LaunchCtx& lc, const float* v, float* out, int64_t n, int rank) {
    int64_t tot = 1; for (int a = 0; a < rank; ++a) tot *= n;
    fill(lc, out, tot, 0.f);
    if (n <= 0) return;
    k_diag_embed<<<grid_for(lc, n), kThreads, 0, lc.stream>>>(v, out, n, diag_step(n, rank)); count(lc);
}
void diag_extract(const LaunchCtx& lc, const float* a, float* out, int64_t n, int rank) {
    if (n <= 0) return;
    k_diag_extract<<<grid_for(lc, n), kThreads, 0, lc.stream>>>(a, out, n, diag_step(n, rank)); count(lc);
}

void softmax_rows(const LaunchCtx& lc, const float* Z, float* A, int64_t rows, int64_t cols) {
    if (rows * cols <= 0) return;
    if (cols <= 16) {
        const int vec = cols % 4 == 0 && aligned16(Z) && aligned16(A);
        k_softmax_small<16, 0><<<grid_for(lc, rows), kThreads, 0, lc.stream>>>(Z, nullptr, A, rows, (int)cols, vec); count(lc);
        return;
    }
    k_softmax_rows<0><<<grid_for(lc, rows * 32), kThreads, 0, lc.stream>>>(Z, nullptr, A, nullptr, nullptr, rows, cols, nullptr); count(lc);
}
void softmax_vjp_rows(const LaunchCtx& lc, const float* Z, const float* dA, float* dZ, int64_t rows, int64_t cols) {
    if (rows * cols <= 0) return;
    if (cols <= 16) {
        const int vec = cols % 4 == 0 && aligned16(Z) && aligned16(dA) && aligned16(dZ);
        k_softmax_small<16, 1><<<grid_for(lc, rows), kThreads, 0, lc.stream>>>(Z, dA, dZ, rows, (int)cols, vec); count(lc);
        return;
    }
    k_softmax_rows<1><<<grid_for(lc, rows * 32), kThreads, 0, lc.stream>>>(Z, dA, nullptr, dZ, nullptr, rows, cols, nullptr); count(lc);
}
bool softmax_ce_rows(const LaunchCtx& lc, const float* Z, const float* Y, float* A, float* dZ, float* loss, int64_t rows, int64_t cols, float* db) {
    if (rows * cols <= 0) return false;
    const bool fuse_db = db != nullptr && cols <= 32;
    if (cols <= 16) {
        k_softmax_ce_small<16><<<grid_for(lc, rows, kThreads, 2), kThreads, 0, lc.stream>>>(Z, Y, A, dZ, loss, rows, (int)cols, fuse_db ? db : nullptr); count(lc);
        return fuse_db;
    }
    k_softmax_rows<2><<<grid_for(lc, rows * 32, kThreads, 4), kThreads, 0, lc.stream>>>(Z, Y, A, dZ, loss, rows, cols, fuse_db ? db : nullptr); count(lc);
    return fuse_db;
}
void loss_vjp(const LaunchCtx& lc, int loss, const float* A, const float* Y, float* dA, float* loss_out, int64_t n) {
    if (n <= 0) return;
    k_loss_vjp<<<grid_for(lc, n, kThreads, 4), kThreads, 0, lc.stream>>>(loss, A, Y, dA, loss_out, n); count(lc);
}

int gemm_simt(const LaunchCtx& lc, const GemmCall& c) {
    if (c.dtype != 0 || c.io_bf16) return -1;
    if (c.M <= 0 || c.N <= 0) return 0;
    GemmParams p{};
    p.M = c.M; p.N = c.N; p.K = c.K;
    p.epi = c.epi; p.act = c.act; p.alpha = c.alpha; p.beta = c.beta;
    p.out0 = c.out0; p.ld_out0 = c.ld_out0; p.out1 = c.out1; p.ld_out1 = c.ld_out1;
    p.aux0 = c.aux0; p.ld_aux0 = c.ld_aux0; p.bias = c.bias; p.loss = c.loss;
    const int gx = (c.N + 63) / 64, gy = (c.M + 63) / 64;
    int gz = 1;
    if (c.epi == EPI_ATOMIC) {
        gz = c.split_k > 0 ? c.split_k : (int)((2LL * lc.num_sms + (long long)gx * gy - 1) / ((long long)gx * gy));
        const int maxz = (c.K + 255) / 256;
        if (gz > maxz) gz = maxz;
        if (gz < 1) gz = 1;
    }
    dim3 g(gx, gy, gz);
    const float* A = reinterpret_cast<const float*>(c.A);
    const float* B = reinterpret_cast<const float*>(c.B);
    if (c.major_a == MAJOR_K && c.major_b == MAJOR_K) k_gemm_simt<MAJOR_K, MAJOR_K><<<g, 256, 0, lc.stream>>>(A, c.lda, B, c.ldb, p);
    else if (c.major_a == MAJOR_K) k_gemm_simt<MAJOR_K, MAJOR_MN><<<g, 256, 0, lc.stream>>>(A, c.lda, B, c.ldb, p);
    else if (c.major_b == MAJOR_K) k_gemm_simt<MAJOR_MN, MAJOR_K><<<g, 256, 0, lc.stream>>>(A, c.lda, B, c.ldb, p);
    else k_gemm_simt<MAJOR_MN, MAJOR_MN><<<g, 256, 0, lc.stream>>>(A, c.lda, B, c.ldb, p);
    count(lc);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}


// ---------------------------------------------------------------------- skinny GEMMs: one of M, N, K is <= 16
// The MLP's output layer (10 classes at BASELINE configs[2]) gives three products with one tiny dimension: 5 FLOP per byte of the
// large operand, i.e. HBM-bound by a factor > 20 on this machine.  A 128 x 256 tensor-core tile would be 4-6 % full and the fp16
// pair splits of the operands would move more bytes than the product itself, so these run on the CUDA cores in plain fp32 (exact
// products, fp32 accumulation) at the speed the large operand streams:
//   skinny-K  out[s,f]  = epi( sum_{j<n} T[s,j] W[j,f] )        (dA of the output layer: EPI_MUL_DACT, + db / max|out| side outputs)
//   skinny-M  out[j,f] += alpha sum_s T[s,j] F[s,f]             (dW of the output layer: EPI_ATOMIC)
//   skinny-N  out[s,j]  = epi( sum_k X[s,k] W[j,k] )            (forward of the output layer: EPI_BIAS_ACT)
namespace {
constexpr int kSkinnyRows = 256;    // rows of the thin operand staged per pass (padded to 16 floats each: 16 KiB)
constexpr int kSkinnyThreads = 512;
constexpr int kSkinnyBatch = 4;     // rows per batch; the next batch's 16-byte loads of the wide operand are in flight while this one is used

__device__ __forceinline__ void block_absmax_commit(float amax, unsigned int* out) {
    if (out == nullptr) return;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(amax));
}

// thin operand rows [r0, r0 + nrows) -> sT[row][16] (zero padded), coalesced
__device__ __forceinline__ void skinny_stage_thin(float (*sT)[16], const float* __restrict__ T, long long ldt, int64_t r0, int nrows, int n) {
    const int tid = threadIdx.y * blockDim.x + threadIdx.x;
    constexpr int PER = kSkinnyRows * 16 / kSkinnyThreads;
    float tmp[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {                  // all loads in flight before the first store
        const int i = tid + u * kSkinnyThreads, r = i >> 4, j = i & 15;
        tmp[u] = (r < nrows && j < n) ? __ldg(T + (r0 + r) * ldt + j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) { const int i = tid + u * kSkinnyThreads; sT[i >> 4][i & 15] = tmp[u]; }
}
template <int VEC> struct SkinnyVec { float v[VEC]; };
template <int VEC> __device__ __forceinline__ SkinnyVec<VEC> skinny_ld(const float* p) {
    SkinnyVec<VEC> r;
    if constexpr (VEC == 4) { const float4 t = __ldg(reinterpret_cast<const float4*>(p)); r.v[0] = t.x; r.v[1] = t.y; r.v[2] = t.z; r.v[3] = t.w; }
    else r.v[0] = __ldg(p);
    return r;
}
template <int VEC> __device__ __forceinline__ void skinny_st(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else *p = v[0];
}

// blockDim = (TX column threads, TY row lanes), 512 threads; thread = VEC consecutive columns; NJ = n rounded up to a multiple of 4.
// Row lane y takes rows y, y + TY, ... of a staged pass, kSkinnyBatch at a time: the batch's loads of the wide operand are issued
// before any of them is used.
template <int VEC, int NJ, bool LOGI>   // LOGI: EPI_MUL_DACT with the logistic derivative a (1 - a), resolved at compile time
__global__ void __launch_bounds__(kSkinnyThreads) k_skinny_k(const float* __restrict__ T, long long ldt, const float* __restrict__ W, long long ldw, GemmParams p, int rows_per_cta) {
    __shared__ __align__(16) float sT[kSkinnyRows][16];
    __shared__ float sCol[kSkinnyThreads * VEC];     // column-sum partials of the row lanes, reduced before the reds
    const int f0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const bool col_ok = f0 < p.N;            // VEC == 4: N % 4 == 0, so a thread is all in or all out
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_cta, r_end = min((int64_t)p.M, r_begin + rows_per_cta);
    const float* aux = reinterpret_cast<const float*>(p.aux0);
    float* out = reinterpret_cast<float*>(p.out0);
    const bool dact = p.epi == EPI_MUL_DACT;
    const int TY = blockDim.y;
    // the wide operand's loads run one batch ahead of the arithmetic (rows r_begin + y + i * TY of the CTA's slice, in slice order)
    const int64_t stride = (int64_t)kSkinnyBatch * TY;
    SkinnyVec<VEC> nxt[kSkinnyBatch];
    auto fetch = [&](int64_t row0) {
        if (!dact || !col_ok) return;
#pragma unroll
        for (int i = 0; i < kSkinnyBatch; ++i)
            if (row0 + i * TY < r_end) nxt[i] = skinny_ld<VEC>(aux + (row0 + i * TY) * p.ld_aux0 + f0);
    };
    fetch(r_begin + threadIdx.y);
    float w[NJ][VEC];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) w[j][v] = (col_ok && j < p.K) ? __ldg(W + (long long)j * ldw + f0 + v) : 0.f;
    float csum[VEC], amax = 0.f;
#pragma unroll
    for (int v = 0; v < VEC; ++v) csum[v] = 0.f;
    static_assert(kSkinnyRows % kSkinnyBatch == 0, "a staged pass holds whole batches");
    for (int64_t r0 = r_begin; r0 < r_end; r0 += kSkinnyRows) {
        const int nrows = (int)min((int64_t)kSkinnyRows, r_end - r0);
        __syncthreads();
        skinny_stage_thin(sT, T, ldt, r0, nrows, p.K);
        __syncthreads();
        if (!col_ok) continue;
        // kSkinnyRows is a multiple of kSkinnyBatch * TY (TY | 16), so the batches of consecutive passes continue the same sequence
        for (int rb = threadIdx.y; rb < nrows; rb += (int)stride) {
            SkinnyVec<VEC> a[kSkinnyBatch];
            int ridx[kSkinnyBatch]; bool okr[kSkinnyBatch];
#pragma unroll
            for (int i = 0; i < kSkinnyBatch; ++i) { a[i] = nxt[i]; okr[i] = rb + i * TY < nrows; ridx[i] = okr[i] ? rb + i * TY : rb; }
            fetch(r0 + rb + stride);
            // the batch's rows are computed together, branch-free (rows past the end repeat row rb and are not stored):
            // kSkinnyBatch * VEC independent FMA chains instead of VEC
            float acc[kSkinnyBatch][VEC];
#pragma unroll
            for (int i = 0; i < kSkinnyBatch; ++i)
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[i][v] = 0.f;
#pragma unroll
            for (int j4 = 0; j4 < NJ / 4; ++j4) {
#pragma unroll
                for (int i = 0; i < kSkinnyBatch; ++i) {
                    const float4 t = *reinterpret_cast<const float4*>(&sT[ridx[i]][j4 * 4]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        acc[i][v] = fmaf(t.x, w[j4 * 4][v], acc[i][v]); acc[i][v] = fmaf(t.y, w[j4 * 4 + 1][v], acc[i][v]);
                        acc[i][v] = fmaf(t.z, w[j4 * 4 + 2][v], acc[i][v]); acc[i][v] = fmaf(t.w, w[j4 * 4 + 3][v], acc[i][v]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < kSkinnyBatch; ++i) {
                if (!okr[i]) continue;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if constexpr (LOGI) acc[i][v] *= a[i].v[v] * (1.0f - a[i].v[v]);
                    else acc[i][v] = dact ? acc[i][v] * act_deriv_from_out(p.act, a[i].v[v]) : acc[i][v] * p.alpha;
                    csum[v] += acc[i][v]; amax = fmaxf(amax, fabsf(acc[i][v]));
                }
                skinny_st<VEC>(out + (r0 + ridx[i]) * p.ld_out0 + f0, acc[i]);
            }
        }
    }
    if (p.colsum != nullptr) {               // uniform across the block
        __syncthreads();
#pragma unroll
        for (int v = 0; v < VEC; ++v) sCol[(threadIdx.y * blockDim.x + threadIdx.x) * VEC + v] = csum[v];
        __syncthreads();
        if (threadIdx.y == 0 && col_ok) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float t = 0.f;
                for (int y = 0; y < TY; ++y) t += sCol[(y * blockDim.x + threadIdx.x) * VEC + v];
                atomicAdd(p.colsum + f0 + v, t);
            }
        }
    }
    block_absmax_commit(amax, p.absmax_out);
}

template <int VEC, int NJ>
__global__ void __launch_bounds__(kSkinnyThreads) k_skinny_m(const float* __restrict__ T, long long ldt, const float* __restrict__ F, long long ldf, GemmParams p, int rows_per_cta) {
    // one buffer, two lives: the staged thin rows during the sweep, then the row lanes' partial sums (4 thin columns at a time)
    __shared__ __align__(16) float sbuf[4 * kSkinnyThreads * VEC > kSkinnyRows * 16 ? 4 * kSkinnyThreads * VEC : kSkinnyRows * 16];
    float (*sT)[16] = reinterpret_cast<float (*)[16]>(sbuf);
    const int f0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
    const bool col_ok = f0 < p.N;
    const int TY = blockDim.y;
    const int64_t r_begin = (int64_t)blockIdx.y * rows_per_cta, r_end = min((int64_t)p.K, r_begin + rows_per_cta);   // the contraction runs over rows
    const int64_t stride = (int64_t)kSkinnyBatch * TY;
    SkinnyVec<VEC> nxt[kSkinnyBatch];
    auto fetch = [&](int64_t row0) {
        if (!col_ok) return;
#pragma unroll
        for (int i = 0; i < kSkinnyBatch; ++i)
            if (row0 + i * TY < r_end) nxt[i] = skinny_ld<VEC>(F + (row0 + i * TY) * ldf + f0);
    };
    fetch(r_begin + threadIdx.y);
    float acc[NJ][VEC];
#pragma unroll
    for (int j = 0; j < NJ; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[j][v] = 0.f;
    for (int64_t r0 = r_begin; r0 < r_end; r0 += kSkinnyRows) {
        const int nrows = (int)min((int64_t)kSkinnyRows, r_end - r0);
        __syncthreads();
        skinny_stage_thin(sT, T, ldt, r0, nrows, p.M);
        __syncthreads();
        if (!col_ok) continue;
        for (int rb = threadIdx.y; rb < nrows; rb += (int)stride) {
            SkinnyVec<VEC> x[kSkinnyBatch];
            int ridx[kSkinnyBatch];
#pragma unroll
            for (int i = 0; i < kSkinnyBatch; ++i) {
                const bool okr = rb + i * TY < nrows;
                ridx[i] = okr ? rb + i * TY : rb;
#pragma unroll
                for (int v = 0; v < VEC; ++v) x[i].v[v] = okr ? nxt[i].v[v] : 0.f;   // rows past the end contribute nothing
            }
            fetch(r0 + rb + stride);
#pragma unroll
            for (int j4 = 0; j4 < NJ / 4; ++j4) {
#pragma unroll
                for (int i = 0; i < kSkinnyBatch; ++i) {
                    const float4 t = *reinterpret_cast<const float4*>(&sT[ridx[i]][j4 * 4]);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        acc[j4 * 4][v] = fmaf(t.x, x[i].v[v], acc[j4 * 4][v]); acc[j4 * 4 + 1][v] = fmaf(t.y, x[i].v[v], acc[j4 * 4 + 1][v]);
                        acc[j4 * 4 + 2][v] = fmaf(t.z, x[i].v[v], acc[j4 * 4 + 2][v]); acc[j4 * 4 + 3][v] = fmaf(t.w, x[i].v[v], acc[j4 * 4 + 3][v]);
                    }
                }
            }
        }
    }
    float* out = reinterpret_cast<float*>(p.out0);
    const int slot = (threadIdx.y * blockDim.x + threadIdx.x) * VEC;
#pragma unroll
    for (int j4 = 0; j4 < NJ / 4; ++j4) {
        if (j4 * 4 >= p.M) break;             // uniform
        __syncthreads();
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
            for (int v = 0; v < VEC; ++v) sbuf[jj * kSkinnyThreads * VEC + slot + v] = acc[j4 * 4 + jj][v];
        __syncthreads();
        // 4 thin columns x (TX * VEC) wide columns to finish: spread over all row lanes (lane y takes thin column y % 4 ... )
        for (int jj = threadIdx.y; jj < 4; jj += TY) {
            const int j = j4 * 4 + jj;
            if (j < p.M && col_ok) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float t = 0.f;
                    for (int y = 0; y < TY; ++y) t += sbuf[jj * kSkinnyThreads * VEC + (y * blockDim.x + threadIdx.x) * VEC + v];
                    atomicAdd(out + (long long)j * p.ld_out0 + f0 + v, p.alpha * t);
                }
            }
        }
    }
}

// halving exchange over the 8 lanes that share a row: lanes whose bit H/2... see k_skinny_n
template <int H, int BIT> __device__ __forceinline__ void skinny_halve(float (&a)[16], int lane) {
    const bool upper = (lane & BIT) != 0;
#pragma unroll
    for (int j = 0; j < H; ++j) {
        const float send = upper ? a[j] : a[j + H];
        const float keep = upper ? a[j + H] : a[j];
        a[j] = keep + __shfl_xor_sync(0xffffffffu, send, BIT);
    }
}

// A warp works on 16 rows: lane = (row quarter q = lane / 8, k lane kl = lane % 8) handles rows q, q + 4, q + 8, q + 12 of the
// group and, of every 8 * VEC-wide chunk of K, the VEC elements at kl * VEC — so a warp load instruction reads four 128-byte row
// segments and the matching pieces of W (shared memory, 128 contiguous bytes broadcast to the four quarters) serve four rows per
// read.  The 8 partial sums of a row meet in a halving exchange (14 shuffles per 16 columns); lane kl ends with columns 2kl, 2kl+1.
template <int VEC, int NJ>
__global__ void __launch_bounds__(256, 2) k_skinny_n(const float* __restrict__ X, long long ldx, const float* __restrict__ W, long long ldw, GemmParams p) {
    extern __shared__ __align__(16) float sW[];
    const int ldk = (p.K + 3) / 4 * 4 + 4;
    for (int j = 0; j < NJ; ++j)                                     // rows N..NJ-1 are zero: the inner loops run over all NJ
        for (int k = threadIdx.x; k < ldk; k += blockDim.x) sW[j * ldk + k] = (j < p.N && k < p.K) ? __ldg(W + (long long)j * ldw + k) : 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31, q = lane >> 3, kl = lane & 7;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float* out = reinterpret_cast<float*>(p.out0);
    const int j0 = 2 * kl;
    const bool biased = p.epi == EPI_BIAS_ACT;
    const float b0 = (biased && p.bias != nullptr && j0 < p.N) ? __ldg(p.bias + j0) : 0.f;
    const float b1 = (biased && p.bias != nullptr && j0 + 1 < p.N) ? __ldg(p.bias + j0 + 1) : 0.f;
    float amax = 0.f;
    constexpr int R = 4, CH = 8 * VEC;
    for (int64_t s0 = warp * 16; s0 < p.M; s0 += nwarps * 16) {
        float acc[R][16];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[r][j] = 0.f;
        const float* xrow[R];
        bool ok[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {   // rows past the end re-read the last row (never stored): the loads need no predicate
            const int64_t row = s0 + q + 4 * r;
            ok[r] = row < p.M;
            xrow[r] = X + (ok[r] ? row : (int64_t)p.M - 1) * ldx + kl * VEC;
        }
        auto chunk = [&](int k0, auto tail_tag) {
            constexpr bool TAIL = decltype(tail_tag)::value;
            const bool kin = !TAIL || k0 + kl * VEC < p.K;
            SkinnyVec<VEC> x[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (kin) x[r] = skinny_ld<VEC>(xrow[r] + k0);
                else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[r].v[v] = 0.f;
                }
            }
            const float* wp = sW + k0 + kl * VEC;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                float wv[VEC];
                if constexpr (VEC == 4) {   // (a tail chunk ends inside the row: lanes beyond it read nothing)
                    const float4 t = kin ? *reinterpret_cast<const float4*>(wp + j * ldk) : make_float4(0.f, 0.f, 0.f, 0.f);
                    wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
                } else wv[0] = kin ? wp[j * ldk] : 0.f;
#pragma unroll
                for (int r = 0; r < R; ++r)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[r][j] = fmaf(x[r].v[v], wv[v], acc[r][j]);
            }
        };
        const int k_main = p.K / CH * CH;
#pragma unroll 2
        for (int k0 = 0; k0 < k_main; k0 += CH) chunk(k0, std::false_type{});
        if (k_main < p.K) chunk(k_main, std::true_type{});
#pragma unroll
        for (int r = 0; r < R; ++r) {
            skinny_halve<8, 4>(acc[r], lane); skinny_halve<4, 2>(acc[r], lane); skinny_halve<2, 1>(acc[r], lane);
            float v0 = acc[r][0], v1 = acc[r][1];
            if (biased) { v0 = act_apply(p.act, v0 + b0); v1 = act_apply(p.act, v1 + b1); }
            else { v0 *= p.alpha; v1 *= p.alpha; }
            if (ok[r]) {
                float* o = out + (s0 + q + 4 * r) * p.ld_out0 + j0;
                if (j0 < p.N) { o[0] = v0; amax = fmaxf(amax, fabsf(v0)); }
                if (j0 + 1 < p.N) { o[1] = v1; amax = fmaxf(amax, fabsf(v1)); }
            }
        }
    }
    block_absmax_commit(amax, p.absmax_out);
}
}  // namespace

size_t skinny_n_smem(int N, int K) { return sizeof(float) * (size_t)((N + 3) / 4 * 4) * ((size_t)(K + 3) / 4 * 4 + 4); }
constexpr size_t kSkinnyNSmemMax = 200 * 1024;   // K <= 3196

// 0 = not a skinny product, 1 = skinny-K, 2 = skinny-M, 3 = skinny-N
int gemm_skinny_kind(const GemmCall& c) {
    if (c.dtype != 0 || c.io_bf16 || c.M <= 0 || c.N <= 0 || c.K <= 0 || c.out0_mc != nullptr) return 0;
    if (c.K <= 16 && c.major_a == MAJOR_K && c.major_b == MAJOR_MN && ((c.epi == EPI_STORE && c.aux0 == nullptr) || c.epi == EPI_MUL_DACT) &&
        (c.colsum == nullptr || c.colsum_src == 0 || c.colsum_src == 1)) return 1;
    if (c.M <= 16 && c.major_a == MAJOR_MN && c.major_b == MAJOR_MN && c.epi == EPI_ATOMIC) return 2;
    if (c.N <= 16 && c.major_a == MAJOR_K && c.major_b == MAJOR_K && ((c.epi == EPI_STORE && c.aux0 == nullptr) || c.epi == EPI_BIAS_ACT) &&
        (c.colsum == nullptr || c.colsum_src == 0) && skinny_n_smem(c.N, c.K) <= kSkinnyNSmemMax) return 3;
    return 0;
}

int gemm_skinny(const LaunchCtx& lc, const GemmCall& c, int* colsum_fused, int* absmax_done) {
    const int kind = gemm_skinny_kind(c);
    if (kind == 0) return 0;
    const float* A = reinterpret_cast<const float*>(c.A);
    const float* B = reinterpret_cast<const float*>(c.B);
    GemmParams p{};
    p.M = c.M; p.N = c.N; p.K = c.K; p.epi = c.epi; p.act = c.act; p.alpha = c.alpha;
    p.out0 = c.out0; p.ld_out0 = c.ld_out0; p.aux0 = c.aux0; p.ld_aux0 = c.ld_aux0; p.bias = c.bias;
    auto a16 = [](const void* q, long long ld) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0 && ld % 4 == 0; };
    // columns of the wide operand per thread / threads per row / row lanes, and the row slices of a grid of ~4 CTAs per SM
    auto shape = [&](int ncols, int64_t rows, bool vec, int ctas_per_sm, dim3& block, dim3& grid, int& rows_per_cta) {
        const int per = vec ? 4 : 1;
        int tx = (ncols + per - 1) / per;
        tx = tx >= kSkinnyThreads ? kSkinnyThreads : (tx + 31) / 32 * 32;
        while (kSkinnyThreads % tx) tx += 32;
        const int ty = kSkinnyThreads / tx;
        block = dim3(tx, ty);
        const int gx = (ncols + tx * per - 1) / (tx * per);
        int64_t gy = ((int64_t)lc.num_sms * ctas_per_sm + gx - 1) / gx;
        const int64_t max_gy = (rows + 4 * ty - 1) / (4 * ty);          // at least 4 rows per row lane
        if (gy > max_gy) gy = max_gy;
        if (gy < 1) gy = 1;
        if (gy > 65535) gy = 65535;
        rows_per_cta = (int)((rows + gy - 1) / gy);
        grid = dim3(gx, (unsigned)((rows + rows_per_cta - 1) / rows_per_cta));
    };
    if (kind == 1) {
        const bool vec = c.N % 4 == 0 && a16(c.B, c.ldb) && a16(c.out0, c.ld_out0) && (c.epi != EPI_MUL_DACT || a16(c.aux0, c.ld_aux0));
        p.colsum = c.colsum_src == 1 ? c.colsum : nullptr; p.colsum_src = c.colsum_src;
        p.absmax_out = c.absmax_out;
        dim3 block, grid; int rpc;
        shape(c.N, c.M, vec, 1, block, grid, rpc);
        const int nj = (c.K + 3) / 4;
#define TOPS_SKINNY_K(V, L) \
        switch (nj) { case 1: k_skinny_k<V, 4, L><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; \
                      case 2: k_skinny_k<V, 8, L><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; \
                      case 3: k_skinny_k<V, 12, L><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; \
                      default: k_skinny_k<V, 16, L><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; }
        const bool logi = c.epi == EPI_MUL_DACT && c.act == ACT_LOGISTIC;
        if (vec) { if (logi) { TOPS_SKINNY_K(4, true) } else { TOPS_SKINNY_K(4, false) } }
        else { if (logi) { TOPS_SKINNY_K(1, true) } else { TOPS_SKINNY_K(1, false) } }
#undef TOPS_SKINNY_K
        count(lc);
        if (colsum_fused) *colsum_fused = p.colsum != nullptr ? 1 : 0;
        if (absmax_done) *absmax_done = p.absmax_out != nullptr ? 1 : 0;
    } else if (kind == 2) {
        const bool vec = c.N % 4 == 0 && a16(c.B, c.ldb);
        dim3 block, grid; int rpc;
        shape(c.N, c.K, vec, 1, block, grid, rpc);   // few CTAs: every one ends with M x columns reds
        const int nj = (c.M + 3) / 4;
#define TOPS_SKINNY_M(V) \
        switch (nj) { case 1: k_skinny_m<V, 4><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; \
                      case 2: k_skinny_m<V, 8><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; \
                      case 3: k_skinny_m<V, 12><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; \
                      default: k_skinny_m<V, 16><<<grid, block, 0, lc.stream>>>(A, c.lda, B, c.ldb, p, rpc); break; }
        if (vec) { TOPS_SKINNY_M(4) } else { TOPS_SKINNY_M(1) }
#undef TOPS_SKINNY_M
        count(lc);
    } else {
        const bool vec = c.K % 4 == 0 && a16(c.A, c.lda);
        p.absmax_out = c.absmax_out;
        const size_t smem = skinny_n_smem(c.N, c.K);
        const int grid = grid_for(lc, ((int64_t)c.M + 15) / 16 * 32, kThreads, 4);
        auto launch = [&](auto kern) -> cudaError_t {
            if (smem > 48 * 1024) {   // opt in (per device, idempotent; cheap next to a product with K > 700)
                cudaError_t e1 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSkinnyNSmemMax);
                if (e1 != cudaSuccess) return e1;
            }
            kern<<<grid, kThreads, smem, lc.stream>>>(A, c.lda, B, c.ldb, p);
            return cudaSuccess;
        };
        cudaError_t e1;
        const int nj = (c.N + 3) / 4;
        if (vec) e1 = nj == 1 ? launch(k_skinny_n<4, 4>) : nj == 2 ? launch(k_skinny_n<4, 8>) : nj == 3 ? launch(k_skinny_n<4, 12>) : launch(k_skinny_n<4, 16>);
        else e1 = nj == 1 ? launch(k_skinny_n<1, 4>) : nj == 2 ? launch(k_skinny_n<1, 8>) : nj == 3 ? launch(k_skinny_n<1, 12>) : launch(k_skinny_n<1, 16>);
        if (e1 != cudaSuccess) return -(int)e1;
        count(lc);
        if (absmax_done) *absmax_done = p.absmax_out != nullptr ? 1 : 0;
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 1 : -(int)e;
}

}  // namespace k
}  // namespace tops
