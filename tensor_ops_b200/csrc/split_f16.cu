// Operand preparation for the F16X3 GEMM mode (TOPS_PREC_F16X3): an fp32 tensor x is stored in HBM as an fp16 PAIR
//     t = x * s,   hi = fp16(t),   lo = fp16(t - hi)          (t - hi is exact in fp32; hi + lo carries 22 bits of t)
// where s is a power of two that moves the largest |x| (of the tensor, or of each row) to [2^13, 2^14): fp16 then keeps
// 11 bits in hi and the next 11 in lo for every element within 2^-18 of that maximum, and its absolute error floor (2^-25 in
// scaled units) is 2^-38 of the maximum — far below the 1e-5 parity bar in any norm-wise comparison.  The products
//     hi*hi + lo*hi + hi*lo        (three kind::f16 tcgen05 passes, fp32 accumulate)
// then reproduce x*y to ~2^-22 relative, at 1.5x the cost of a TF32 pass (fp16 MMAs run at twice the TF32 rate).
// All kernels here are HBM-bound streaming passes: 16-byte loads, 8/16-byte stores, grids in multiples of the SM count.
#include <cuda_fp16.h>

#include "kernels.h"

namespace tops {
namespace k {

namespace {

constexpr int kThreads = 256;
inline void count(const LaunchCtx& lc) { if (lc.launches) ++*lc.launches; }

// power-of-two scale that moves m into [2^13, 2^14); returned as the exponent e of m (scale = 2^(13-e)), 0/inf/nan -> scale 1
__device__ __forceinline__ int scale_exp_of(unsigned abs_bits) {
    const int ef = (int)((abs_bits >> 23) & 0xffu);
    if (abs_bits == 0u || ef == 0xff) return 13;          // all-zero or non-finite data: scale 2^0
    int e = ef - 127;
    return e < -100 ? -100 : (e > 100 ? 100 : e);           // keep 2^(13-e) and its inverse normal fp32 numbers
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float((unsigned)(127 + e) << 23); }

// A small second tensor (the layer's W) that rides along with the row split of X instead of taking two launches of its own:
// its max|.| is reduced by the row-split launch, its fp16 pair is written by the fix-up launch.
struct SideSplit { const float* w; int64_t n; unsigned* mx; __half* hi; __half* lo; float* scale2; };

__device__ __forceinline__ unsigned warp_max_u(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ void split4(const float4 v, float s, uint2& hi, uint2& lo) {
    const float t0 = v.x * s, t1 = v.y * s, t2 = v.z * s, t3 = v.w * s;
    __half2 h0 = __floats2half2_rn(t0, t1), h1 = __floats2half2_rn(t2, t3);
    const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
    __half2 l0 = __floats2half2_rn(t0 - f0.x, t1 - f0.y), l1 = __floats2half2_rn(t2 - f1.x, t3 - f1.y);
    hi = make_uint2(*reinterpret_cast<unsigned*>(&h0), *reinterpret_cast<unsigned*>(&h1));
    lo = make_uint2(*reinterpret_cast<unsigned*>(&l0), *reinterpret_cast<unsigned*>(&l1));
}
__device__ __forceinline__ void split1(float v, float s, __half& hi, __half& lo) {
    const float t = v * s;
    hi = __float2half_rn(t);
    lo = __float2half_rn(t - __half2float(hi));
}

// max |x| over a tensor as an fp32 bit pattern (monotonic for non-negative floats): atomicMax into a pre-zeroed word
__global__ void __launch_bounds__(kThreads) k_absmax(const float* __restrict__ x, int64_t n, unsigned* __restrict__ out, int vec) {
    unsigned m = 0u;
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const int64_t n4 = n >> 2;
        int64_t i = tid;
        for (; i + 3 * stride < n4; i += 4 * stride) {     // four independent 16-byte loads in flight per thread
            const float4 a = __ldg(x4 + i), b = __ldg(x4 + i + stride), c = __ldg(x4 + i + 2 * stride), d = __ldg(x4 + i + 3 * stride);
            const float ma = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
            const float mb = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
            const float mc = fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w)));
            const float md = fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)));
            m = max(m, __float_as_uint(fmaxf(fmaxf(ma, mb), fmaxf(mc, md))));
        }
        for (; i < n4; i += stride) {
            const float4 a = __ldg(x4 + i);
            m = max(m, __float_as_uint(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)))));
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += stride) m = max(m, __float_as_uint(fabsf(x[j])));
    } else {
        for (int64_t i = tid; i < n; i += stride) m = max(m, __float_as_uint(fabsf(x[i])));
    }
    // NaN inputs: fmaxf drops them, so the scale stays finite; the NaN itself still reaches the fp16 planes and the result
    m = warp_max_u(m);
    __shared__ unsigned sh[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < kThreads / 32 ? sh[threadIdx.x] : 0u;
        m = warp_max_u(m);
        if (threadIdx.x == 0 && m != 0u) atomicMax(out, m);
    }
}

// whole-tensor split with one scale derived from absmax bits; thread 0 also publishes {scale, 1/scale}
__global__ void __launch_bounds__(kThreads) k_split_f16_tensor(const float* __restrict__ x, int64_t n, const unsigned* __restrict__ absmax_bits,
                                                               __half* __restrict__ hi, __half* __restrict__ lo, float* __restrict__ scale2, int vec) {
    const int e = scale_exp_of(__ldg(absmax_bits));
    const float s = pow2f(13 - e);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0 && scale2 != nullptr) { scale2[0] = s; scale2[1] = pow2f(e - 13); }
    if (vec) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        uint2* h2 = reinterpret_cast<uint2*>(hi);
        uint2* l2 = reinterpret_cast<uint2*>(lo);
        const int64_t n4 = n >> 2;
        for (int64_t i = tid; i < n4; i += stride) {
            uint2 h, l;
            split4(__ldg(x4 + i), s, h, l);
            h2[i] = h; l2[i] = l;
        }
        for (int64_t j = (n4 << 2) + tid; j < n; j += stride) split1(x[j], s, hi[j], lo[j]);
    } else {
        for (int64_t i = tid; i < n; i += stride) split1(x[i], s, hi[i], lo[i]);
    }
}

// same, for a [rows, cols] matrix whose fp16 planes have a padded leading dimension ld_out >= cols (rows of any length made
// TMA-expressible: 16-byte aligned fp16 rows); the padding elements are never read (the tensor map's inner extent is cols)
__global__ void __launch_bounds__(kThreads) k_split_f16_tensor_2d(const float* __restrict__ x, int64_t rows, int64_t cols, int64_t ld_out,
                                                                  const unsigned* __restrict__ absmax_bits, __half* __restrict__ hi, __half* __restrict__ lo,
                                                                  float* __restrict__ scale2) {
    const int e = scale_exp_of(__ldg(absmax_bits));
    const float s = pow2f(13 - e);
    const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
    if (tid == 0 && scale2 != nullptr) { scale2[0] = s; scale2[1] = pow2f(e - 13); }
    const int64_t n = rows * cols;
    for (int64_t i = tid; i < n; i += stride) {
        const int64_t r = i / cols, c = i - r * cols;
        split1(x[i], s, hi[r * ld_out + c], lo[r * ld_out + c]);
    }
}

__global__ void k_f16x3_pair_scale(const float* __restrict__ sa2, const float* __restrict__ sb2, float* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = sa2[1] * sb2[1];   // 1 / (sA sB): powers of two, exact
}

// Row-wise split: one warp per row, the row held in registers (NV float4 per lane) between the max and the conversion so that
// x is read from HBM exactly once.  rs[r] = 1 / scale_r; max_rs (fp32 bits, pre-zeroed) = max_r rs[r].
// Optionally (y != NULL) the same launch also reduces max |y| over a second [rows, ycols] tensor into ymax_bits: the rows of
// the cotangent dA ride along with the rows of X in ffLayer's gradient, saving one launch and its tail.
template <int NV>
__global__ void __launch_bounds__(kThreads) k_split_f16_rows(const float* __restrict__ x, int64_t rows, int64_t cols, __half* __restrict__ hi,
                                                             __half* __restrict__ lo, float* __restrict__ rs, unsigned* __restrict__ max_rs_bits,
                                                             const float* __restrict__ y, int64_t ycols, unsigned* __restrict__ ymax_bits, SideSplit side) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t c4 = cols >> 2;    // cols % 4 == 0 and 16-byte aligned rows are guaranteed by the launcher for NV > 0
    if (side.w != nullptr) {         // max |W| (16-byte aligned, n % 4 == 0: checked by the launcher)
        const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
        const float4* w4 = reinterpret_cast<const float4*>(side.w);
        unsigned m = 0u;
        for (int64_t i = tid; i < (side.n >> 2); i += stride) {
            const float4 a = __ldg(w4 + i);
            m = max(m, __float_as_uint(fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)))));
        }
        m = warp_max_u(m);
        if (lane == 0 && m != 0u) atomicMax(side.mx, m);
    }
    unsigned rs_max = 0u, ymax = 0u;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float4* xr = reinterpret_cast<const float4*>(x + r * cols);
        float4 v[NV];
        float m = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int64_t c = lane + 32 * k;
            v[k] = c < c4 ? __ldg(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(v[k].x), fabsf(v[k].y)), fmaxf(fabsf(v[k].z), fabsf(v[k].w))));
        }
        if (y != nullptr) {          // cotangent row: max only
            const float4* yr = reinterpret_cast<const float4*>(y + r * ycols);
            float my = 0.f;
            for (int64_t c = lane; c < (ycols >> 2); c += 32) {
                const float4 a = __ldg(yr + c);
                my = fmaxf(my, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
            }
            ymax = max(ymax, __float_as_uint(my));
        }
        const unsigned m_all = warp_max_u(__float_as_uint(m));
        const int e = scale_exp_of(m_all);
        const float s = pow2f(13 - e);
        uint2* hr = reinterpret_cast<uint2*>(hi + r * cols);
        uint2* lr = reinterpret_cast<uint2*>(lo + r * cols);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int64_t c = lane + 32 * k;
            if (c < c4) {
                uint2 h, l;
                split4(v[k], s, h, l);
                hr[c] = h; lr[c] = l;
            }
        }
        const float inv = m_all == 0u ? 0.f : pow2f(e - 13);   // all-zero row: marked, its scale is chosen by the fix-up pass
        if (lane == 0) rs[r] = inv;
        rs_max = max(rs_max, __float_as_uint(inv));
    }
    if (lane == 0 && rs_max != 0u) atomicMax(max_rs_bits, rs_max);
    if (y != nullptr) {
        ymax = warp_max_u(ymax);
        if (lane == 0 && ymax != 0u) atomicMax(ymax_bits, ymax);
    }
}

// generic rows (any cols / alignment): two passes over the row, the second one served by L1/L2
__global__ void __launch_bounds__(kThreads) k_split_f16_rows_generic(const float* __restrict__ x, int64_t rows, int64_t cols, __half* __restrict__ hi,
                                                                     __half* __restrict__ lo, float* __restrict__ rs, unsigned* __restrict__ max_rs_bits) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    unsigned rs_max = 0u;
    for (int64_t r = warp; r < rows; r += nwarps) {
        const float* xr = x + r * cols;
        float m = 0.f;
        for (int64_t c = lane; c < cols; c += 32) m = fmaxf(m, fabsf(xr[c]));
        const unsigned m_all = warp_max_u(__float_as_uint(m));
        const int e = scale_exp_of(m_all);
        const float s = pow2f(13 - e);
        for (int64_t c = lane; c < cols; c += 32) split1(xr[c], s, hi[r * cols + c], lo[r * cols + c]);
        const float inv = m_all == 0u ? 0.f : pow2f(e - 13);
        if (lane == 0) rs[r] = inv;
        rs_max = max(rs_max, __float_as_uint(inv));
    }
    if (lane == 0 && rs_max != 0u) atomicMax(max_rs_bits, rs_max);
}

// Fix-up after the row split, once max_s rs[s] is final.  The cotangent pair is scaled by c * rs[s] (so that the per-sample factors
// cancel inside dW's contraction over s), with c = 2^13 / (max|dA| max_s rs[s]): a row whose rs is 2^-k of the largest gets a
// cotangent pair 2^-k below full scale.  Without a bound on k one outlier row (magnitude 1e6 times the others) would push every
// other row's dZ pair into fp16's subnormals — and their dX with it.  So rows more than kRowScaleSpread binades below the largest
// are re-split with the scale of the bound (they keep >= 22 - 8 bits of the largest row's resolution, which is what a per-tensor
// scale would have given them, times 2^8), and all-zero rows (marked rs = 0, planes are zeros under any scale) get the largest rs.
// In the common case every row passes the test and the pass only reads rs[]: a few microseconds.
constexpr int kRowScaleSpread = 8;
__global__ void __launch_bounds__(kThreads) k_split_f16_rows_fixup(const float* __restrict__ x, int64_t rows, int64_t cols, __half* __restrict__ hi,
                                                                   __half* __restrict__ lo, float* __restrict__ rs, const unsigned* __restrict__ max_rs_bits,
                                                                   SideSplit side, const unsigned* __restrict__ absmax_dA_bits, const float* __restrict__ sW2,
                                                                   float* __restrict__ scales_out) {
    const int lane = threadIdx.x & 31;
    {   // riders of this launch: the fp16 pair of the side tensor (its max is final since the previous launch) and the layer's scalars
        const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
        float inv_sW = sW2 != nullptr ? sW2[1] : 1.0f;
        if (side.w != nullptr) {
            const int e = scale_exp_of(__ldg(side.mx));
            const float sw = pow2f(13 - e);
            inv_sW = pow2f(e - 13);
            if (tid == 0 && side.scale2 != nullptr) { side.scale2[0] = sw; side.scale2[1] = inv_sW; }
            const float4* w4 = reinterpret_cast<const float4*>(side.w);
            uint2* h2 = reinterpret_cast<uint2*>(side.hi);
            uint2* l2 = reinterpret_cast<uint2*>(side.lo);
            for (int64_t i = tid; i < (side.n >> 2); i += stride) {
                uint2 h, l;
                split4(__ldg(w4 + i), sw, h, l);
                h2[i] = h; l2[i] = l;
            }
        }
        if (scales_out != nullptr && tid == 0) {          // see k_f16x3_layer_scales
            const int eD = scale_exp_of(*absmax_dA_bits);
            const unsigned rb = *max_rs_bits;
            const int eR = rb == 0u ? 0 : (int)((rb >> 23) & 0xffu) - 127;
            int ec = 13 - eD - eR;
            ec = ec < -120 ? -120 : (ec > 120 ? 120 : ec);
            scales_out[0] = inv_sW; scales_out[1] = pow2f(ec); scales_out[2] = pow2f(-ec); scales_out[3] = pow2f(-ec) * inv_sW;
        }
    }
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned mb = __ldg(max_rs_bits);
    const float mr = mb == 0u ? 1.0f : __uint_as_float(mb);              // every row zero: any scale will do
    const int er = (int)((__float_as_uint(mr) >> 23) & 0xffu) - 127;
    const int et = er - kRowScaleSpread < -120 ? -120 : er - kRowScaleSpread;
    const float thr = pow2f(et), s = pow2f(-et);
    // a warp tests 32 rows with one coalesced load; only rows that fail the test (none, normally) are touched again
    for (int64_t r0 = warp * 32; r0 < rows; r0 += nwarps * 32) {
        const int64_t rl = r0 + lane;
        const float inv = rl < rows ? rs[rl] : mr;
        if (inv == 0.f) rs[rl] = mr;                                      // all-zero row
        unsigned todo = __ballot_sync(0xffffffffu, inv != 0.f && inv < thr);
        while (todo) {
            const int j = __ffs(todo) - 1;
            todo &= todo - 1;
            const int64_t r = r0 + j;
            const float* xr = x + r * cols;
            for (int64_t c = lane; c < cols; c += 32) split1(xr[c], s, hi[r * cols + c], lo[r * cols + c]);
            if (lane == 0) rs[r] = thr;
        }
    }
}

// The scalars one ffLayer forward + VJP needs (all powers of two):
//   out[0] = 1/sW                forward:  Z  = acc * out[0] * rsX[s]
//   out[1] = c                   dZ pair:  dZ' = dZ * c * rsX[s],  c = 2^(13 - e(max|dA|)) / max_s rsX[s]  =>  |dZ'| < 2^14  (|act'| <= 1)
//   out[2] = 1/c                 dW = acc * out[2]          (the per-sample factors rsX[s] * sX[s] cancel inside the contraction)
//   out[3] = 1/(c sW)            dX = acc * out[3] / rsX[s]
__global__ void k_f16x3_layer_scales(const unsigned* __restrict__ max_rs_bits, const unsigned* __restrict__ absmax_dA_bits, const float* __restrict__ sW2,
                                     float* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int eD = scale_exp_of(*absmax_dA_bits);
    const unsigned rb = *max_rs_bits;
    const int eR = rb == 0u ? 0 : (int)((rb >> 23) & 0xffu) - 127;     // max rs is an exact power of two
    int ec = 13 - eD - eR;
    ec = ec < -120 ? -120 : (ec > 120 ? 120 : ec);
    const float inv_sW = sW2 != nullptr ? sW2[1] : 1.0f;
    out[0] = inv_sW;
    out[1] = pow2f(ec);
    out[2] = pow2f(-ec);
    out[3] = pow2f(-ec) * inv_sW;
}

inline int grid_for(const LaunchCtx& lc, int64_t work_items, int per_block, int waves) {
    int64_t b = (work_items + per_block - 1) / per_block;
    const int64_t cap = (int64_t)lc.num_sms * waves;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

void absmax_bits(const LaunchCtx& lc, const float* x, int64_t n, unsigned* out_bits) {
    if (n <= 0) return;
    const int vec = aligned16(x) ? 1 : 0;
    k_absmax<<<grid_for(lc, (n + 3) / 4, kThreads, 8), kThreads, 0, lc.stream>>>(x, n, out_bits, vec);
    count(lc);
}

void split_f16_tensor(const LaunchCtx& lc, const float* x, int64_t n, const unsigned* absmax_bits_dev, void* hi, void* lo, float* scale2) {
    if (n <= 0) return;
    const int vec = (aligned16(x) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7) == 0) ? 1 : 0;
    k_split_f16_tensor<<<grid_for(lc, (n + 3) / 4, kThreads, 8), kThreads, 0, lc.stream>>>(x, n, absmax_bits_dev, (__half*)hi, (__half*)lo, scale2, vec);
    count(lc);
}

void split_f16_tensor_2d(const LaunchCtx& lc, const float* x, int64_t rows, int64_t cols, int64_t ld_out, const unsigned* absmax_bits_dev, void* hi, void* lo,
                         float* scale2) {
    if (rows <= 0 || cols <= 0) return;
    if (ld_out == cols) { split_f16_tensor(lc, x, rows * cols, absmax_bits_dev, hi, lo, scale2); return; }
    k_split_f16_tensor_2d<<<grid_for(lc, rows * cols, kThreads, 8), kThreads, 0, lc.stream>>>(x, rows, cols, ld_out, absmax_bits_dev, (__half*)hi, (__half*)lo, scale2);
    count(lc);
}

void f16x3_pair_scale(const LaunchCtx& lc, const float* sa2, const float* sb2, float* out) {
    k_f16x3_pair_scale<<<1, 32, 0, lc.stream>>>(sa2, sb2, out);
    count(lc);
}

void split_f16_rows(const LaunchCtx& lc, const float* x, int64_t rows, int64_t cols, void* hi, void* lo, float* rs, unsigned* max_rs_bits,
                    const float* y, int64_t ycols, unsigned* ymax_bits, const SideTensor* side_in, const float* sW2, float* scales_out4) {
    if (rows <= 0 || cols <= 0) return;
    const bool fast = (cols % 4) == 0 && aligned16(x) && (reinterpret_cast<uintptr_t>(hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(lo) & 7) == 0 &&
                      cols <= 128 * 16 && (y == nullptr || ((ycols % 4) == 0 && aligned16(y)));
    const int grid = grid_for(lc, rows, kThreads / 32, 8);
    int grid_fix = grid_for(lc, (rows + 31) / 32, kThreads / 32, 2);
    // the side tensor rides along only in its vectorised form; otherwise (and on the generic row path) it takes its own two launches
    SideSplit side{nullptr, 0, nullptr, nullptr, nullptr, nullptr};
    bool side_own = false;
    if (side_in != nullptr && side_in->w != nullptr && side_in->n > 0) {
        if (fast && (side_in->n % 4) == 0 && aligned16(side_in->w) && (reinterpret_cast<uintptr_t>(side_in->hi) & 7) == 0 && (reinterpret_cast<uintptr_t>(side_in->lo) & 7) == 0)
            side = SideSplit{side_in->w, side_in->n, side_in->mx, (__half*)side_in->hi, (__half*)side_in->lo, side_in->scale2};
        else side_own = true;
    }
    if (side.w != nullptr) {   // the rider's 16-byte pieces, one per thread, decide the grid of the fix-up launch
        const int g2 = grid_for(lc, side.n / 4, kThreads, 8);
        if (g2 > grid_fix) grid_fix = g2;
    }
    if (side_own) {
        absmax_bits(lc, side_in->w, side_in->n, side_in->mx);
        split_f16_tensor(lc, side_in->w, side_in->n, side_in->mx, side_in->hi, side_in->lo, side_in->scale2);
        sW2 = side_in->scale2;
    }
    if (!fast) {
        k_split_f16_rows_generic<<<grid, kThreads, 0, lc.stream>>>(x, rows, cols, (__half*)hi, (__half*)lo, rs, max_rs_bits);
        count(lc);
        if (y != nullptr) absmax_bits(lc, y, rows * ycols, ymax_bits);
        k_split_f16_rows_fixup<<<grid_fix, kThreads, 0, lc.stream>>>(x, rows, cols, (__half*)hi, (__half*)lo, rs, max_rs_bits, side, ymax_bits, sW2, scales_out4);
        count(lc);
        return;
    }
    const int64_t c4 = cols / 4;
#define TOPS_SPLIT_ROWS(NV) k_split_f16_rows<NV><<<grid, kThreads, 0, lc.stream>>>(x, rows, cols, (__half*)hi, (__half*)lo, rs, max_rs_bits, y, ycols, ymax_bits, side)
    if (c4 <= 32) TOPS_SPLIT_ROWS(1);
    else if (c4 <= 64) TOPS_SPLIT_ROWS(2);
    else if (c4 <= 128) TOPS_SPLIT_ROWS(4);
    else if (c4 <= 256) TOPS_SPLIT_ROWS(8);
    else TOPS_SPLIT_ROWS(16);
#undef TOPS_SPLIT_ROWS
    count(lc);
    k_split_f16_rows_fixup<<<grid_fix, kThreads, 0, lc.stream>>>(x, rows, cols, (__half*)hi, (__half*)lo, rs, max_rs_bits, side, ymax_bits, sW2, scales_out4);
    count(lc);
}

void f16x3_layer_scales(const LaunchCtx& lc, const unsigned* max_rs_bits, const unsigned* absmax_dA_bits, const float* sW2, float* out4) {
    k_f16x3_layer_scales<<<1, 32, 0, lc.stream>>>(max_rs_bits, absmax_dA_bits, sW2, out4);
    count(lc);
}

}  // namespace k
}  // namespace tops
