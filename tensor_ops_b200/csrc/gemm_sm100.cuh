// Persistent, warp-specialised tcgen05 GEMM for sm_100a.
//
//   C[M,N] = epilogue( sum_k A(m,k) * B(n,k) )
//
// Operands are fp32 (consumed as TF32, one pass, or as a 3-pass hi/lo split that recovers fp32 accuracy)
// or bf16; accumulation is fp32 in TMEM.  Either operand may be K-major (row = MN index, K contiguous) or
// MN-major (row = K index, MN contiguous) so that the three GEMMs of an ffLayer forward + VJP
//   Z  = X  W^T        (A = X  K-major,  B = W  K-major)
//   dX = dZ W          (A = dZ K-major,  B = W  MN-major)
//   dW = dZ^T X        (A = dZ MN-major, B = X  MN-major, split-K)
// run on the tensor cores without any transposed copy in HBM.
//
// One CTA per SM, persistent over work items (output tile x split-K partition).  By default two CTAs form a cluster and work as a
// CTA PAIR on a 256-row tile (CG = 2, tcgen05 cta_group::2): each stages its 128 rows of A and half of the B tile, the leader issues
// the MMAs for both, accumulators stay in each CTA's own TMEM.  CG = 1 is the same code with one CTA per 128-row tile.
//
// Roles (384 threads 1-pass, 512 threads 3-pass; registers rebalanced between warpgroups with setmaxnreg):
//   warp 0      TMA producer      global -> 128B-swizzled smem ring (mbarrier full[])
//   warp 1      MMA issuer        one thread (leader CTA) issues tcgen05.mma into TMEM, commits to empty[] / tmem_full[] (multicast for CG = 2)
//   warp 2      TMEM allocator
//   warp 3      relay (CG = 2, 1-pass): forwards "my stage has landed" to the leader's ready[] barrier
//   warps 4-11  epilogue          2 warps per TMEM lane quarter, each owning half of the tile's columns: tcgen05.ld -> registers -> fused
//                                 bias / logistic / VJP / loss math -> 32x32 staging block in smem -> coalesced 16-byte global stores;
//                                 the aux operand arrives by TMA into a staging block; column sums (db) are taken from the staged block.
//                                 1-pass: the whole K range of a work item accumulates in one TMEM buffer, double-buffered across work items.
//                                 3-pass: the tensor core's accumulator add TRUNCATES (measured: error grows linearly, ~2.7e-8 relative per
//                                 MMA; 8e-6 after K = 1024), so TMEM only ever holds a CHUNK of `chunk_kb` k-blocks; each chunk is drained
//                                 and added, round-to-nearest fp32, into 128 register accumulators per thread while the MMA warp fills the
//                                 other TMEM buffer; the fused epilogue math then runs from the registers.
//   warps 12-15 splitter (3-pass) lo = x - trunc_tf32(x) into a second tile (ready[]); the raw tile serves as hi
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "umma.cuh"

namespace tops {

enum : int { MAJOR_K = 0, MAJOR_MN = 1 };
enum : int {
    EPI_STORE = 0,        // out0 = alpha*acc (+ beta*aux0)
    EPI_ATOMIC = 1,       // out0 += alpha*acc                (split-K partials, red.global.add)
    EPI_BIAS_ACT = 2,     // out0 = act(acc + bias[n])
    EPI_BIAS_ACT_DZ = 3,  // a = act(acc + bias[n]); out0 = a; out1 = aux0 * act'(a)           (aux0 = dA)
    EPI_MUL_DACT = 4,     // out0 = acc * act'(aux0)                                          (aux0 = A of the previous layer)
    EPI_BIAS_ACT_SE = 5,  // a = act(acc+bias); d = aux0 - a; loss += d*d; out0 = a; out1 = -2 d act'(a)   (aux0 = target)
};
enum : int { ACT_ID = 0, ACT_LOGISTIC = 1 };

struct GemmParams {
    int M, N, K;
    int num_m_tiles, num_n_tiles;
    int num_k_blocks;   // ceil(K / KB_ELEMS)
    int split_k;        // number of K partitions (>= 1)
    int kb_per_split;   // k-blocks per partition
    int chunk_kb;       // 3-pass only: k-blocks accumulated in TMEM before promotion to fp32 registers (>= 1)
    int chunk_head_kb;  // chunked kernels: length of the FIRST TWO chunks of a work item (>= chunk_kb).  The two TMEM buffers let the MMA
                        // warp run two chunks into the next work item while the epilogue warps finish the previous one; longer head
                        // chunks buy the fused epilogue that much more time, at the accuracy of a longer TMEM accumulation on them
    int epi, act;
    float alpha, beta;
    void* out0; long long ld_out0;
    void* out1; long long ld_out1;
    const void* aux0; long long ld_aux0;
    const float* bias;
    float* loss;              // scalar accumulator (EPI_BIAS_ACT_SE)
    float* colsum;            // optional [N] accumulator (pre-zeroed): column sums of out0 (colsum_src = 1) or out1 (= 2), TMA epilogue only
    int colsum_src;
    // Fused all-reduce (EPI_ATOMIC only): split-K partials are summed locally in out0; the LAST partial to finish a region
    // (counted in tile_counters) pushes the finished region once to the NVLS multicast alias out0_mc with multimem.red, so the
    // NVSwitch adds it into every rank's replica while the rest of the GEMM is still running.
    float* out0_mc;
    int* tile_counters;       // [num_tiles * CG * 8] zeroed by the caller
    int io_bf16;              // out0/out1/aux0 are bf16 (bf16 pipelines); 0 = fp32
    int vec_ok;               // 16-byte vector access legal for out0/out1/aux0
    int b_presplit;           // PASSES == 2: bf16(B) and bf16(B_lo) exist in HBM (tmB16 / tmBlo16): TMA loads them, the splitter handles A only
    int tma_epi;              // staged epilogue: outputs via smem staging + coalesced stores, aux operand by TMA (tmAux valid if needed)
    // Chunked kernels: the register accumulators are multiplied by (*acc_scale_ptr) * row_scale[m] (or / row_scale[m]) before the
    // epilogue math.  F16X3 operands are stored scaled by powers of two; these factors undo the scaling exactly.
    const float* acc_scale_ptr;   // device scalar, NULL = 1
    const float* acc_scale_ptr2;  // a second device scalar multiplied in (the other operand's inverse scale), NULL = 1
    const float* row_scale;       // [M], NULL = 1
    int row_scale_inv;            // 1: divide by row_scale[m] instead of multiplying
    // out1 written as an fp16 PAIR (staged fp32 epilogues only): t = out1 * (*out1_scale_ptr) * out1_row_scale[m];
    // out1 <- fp16(t), out1b <- fp16(t - fp16(t)); ld_out1 counts fp16 elements.  The column sums (colsum_src = 2) use the fp32 values.
    int out1_pair;
    void* out1b;
    const float* out1_scale_ptr;  // device scalar, NULL = 1
    const float* out1_row_scale;  // [M], NULL = 1
    unsigned int* absmax_out;     // fp16-pair kernels, EV = 1: max over |out0| (float bits, atomicMax; caller zeroes); NULL = none
    unsigned int* watchdog;   // mapped host memory, 2 words
    int debug;                // TOPS_GEMM_DEBUG bit mask (A/B experiments): 1 = no specialised epilogues
};

template <typename T, int MA, int MB, int BN, int STAGES, int PASSES, int CG>
struct GemmCfg {
    static constexpr int BM = 128;
    static constexpr int KB_ELEMS = 128 / (int)sizeof(T);   // K elements per stage (one 128B swizzle row)
    static constexpr int UMMA_K = 32 / (int)sizeof(T);      // K elements per tcgen05.mma
    static constexpr int KSTEPS = KB_ELEMS / UMMA_K;        // = 4
    static constexpr int A_BYTES = BM * 128;
    static constexpr int B_BYTES = (BN / CG) * 128;   // a CTA pair (CG = 2) splits the B tile: each CTA stages N/2 rows
    static_assert(CG == 1 || CG == 2, "CTA group size");
    static constexpr int RAW_BYTES = A_BYTES + B_BYTES;
    // split modes keep a second set of tiles per stage: PASSES == 3 the fp32 lo tiles (RAW_BYTES); PASSES == 2 four bf16 tiles
    // (bf16(A), bf16(A_lo), bf16(B), bf16(B_lo)), half the bytes each — RAW_BYTES in total as well
    // PASSES == 4 (F16X3): fp16 operands pre-split in HBM into (hi, lo) planes, both delivered by TMA — 3 MMA passes
    // hi*hi + lo*hi + hi*lo like PASSES == 3, but no splitter warps and fp16 tensor-core rate
    static constexpr bool PRESPLIT = PASSES == 4;
    static constexpr int NPASS = PRESPLIT ? 3 : PASSES;
    static constexpr int IO_BYTES = PRESPLIT ? 4 : (int)sizeof(T);   // element size of out0 / out1 / aux0
    static constexpr int STAGE_BYTES = RAW_BYTES * (PASSES >= 2 ? 2 : 1);
    static constexpr int NUM_THREADS = (PASSES == 2 || PASSES == 3) ? 512 : 384;
    static constexpr int EPI_THREADS = 256;                 // 8 epilogue warps: 2 per TMEM lane quarter, each owning half of the tile's columns
    static constexpr int TMEM_COLS = 2 * BN;
    static constexpr int BAR_BYTES = 512;
    // epilogue staging: 32x32-element blocks (4 KiB fp32 / 2 KiB bf16) per epilogue warp, moved by TMA.
    // 1-pass: one block for the aux operand (prefetched one chunk ahead) + one for outputs; 3-pass: one shared block.
    static constexpr int EPI_WARPS = EPI_THREADS / 32;
    // columns per epilogue block.  F16X3: 16, so that an aux block AND an output block (2 KiB each) fit beside three 64 KiB stages
    // and the aux operand of sub-block c+1 arrives by TMA while sub-block c is being computed and written
    static constexpr int EPI_W = PRESPLIT ? 16 : 32;
    static constexpr int EPI_BLOCK_BYTES = 32 * EPI_W * IO_BYTES;          // 4096 (fp32, 128-byte rows) / 2048 (bf16, 64-byte rows)
    // two blocks per warp (aux operand prefetched while the previous block is written out) when the pipeline stages leave
    // room for them, otherwise one block shared by the aux operand and the outputs
    static constexpr int EPI_NBUF = (232448 - STAGES * STAGE_BYTES - 2048 >= EPI_WARPS * 2 * EPI_BLOCK_BYTES) ? 2 : 1;
    static constexpr bool EPI_SHARED = EPI_NBUF == 1;
    static constexpr int EPI_WARP_BYTES = EPI_NBUF * EPI_BLOCK_BYTES;
    static constexpr int EPI_BYTES = EPI_WARPS * EPI_WARP_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*barriers*/ + EPI_BYTES + 1024 /*alignment slack*/;
    static_assert(BAR_BYTES <= 1024 && (3 * STAGES + 4 + EPI_WARPS) * 8 + 4 <= BAR_BYTES, "barrier area");
    static_assert(PASSES == 1 || (PASSES >= 2 && PASSES <= 3 && sizeof(T) == 4) || (PASSES == 4 && std::is_same<T, __half>::value),
                  "in-kernel hi/lo split modes take fp32 operands; the pre-split mode takes fp16 pairs");
    static_assert(BN == 128 || BN == 256, "BN");
    static_assert(SMEM_BYTES <= 232448, "shared memory budget");
};

__device__ __forceinline__ float act_apply(int act, float z) {
    if (act == ACT_LOGISTIC) {                                             // 1 / (1 + exp(-z)), NeuralNet.hs:42-44
        float e, r;                                                        // ex2.approx / rcp.approx: 2 MUFU + FMUL + FADD, rel. error ~2e-7
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * -1.4426950408889634f));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
        return r;
    }
    return z;
}
__device__ __forceinline__ float act_deriv_from_out(int act, float a) {
    if (act == ACT_LOGISTIC) return a * (1.0f - a);                        // NeuralNet.hs:46-50
    return 1.0f;
}

// W consecutive elements of one row, registers <-> global (direct epilogue path)
template <bool BF16, int W>
__device__ __forceinline__ void ld_row(const void* base, long long ld, int row, int col, int N, bool vec, float (&x)[W]) {
    if constexpr (BF16) {
        const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(base) + (long long)row * ld + col;
        if (vec && col + W <= N) {
#pragma unroll
            for (int q = 0; q < W / 8; ++q) {
                uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + q);
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 f = __bfloat1622float2(h[e]);
                    x[q * 8 + e * 2] = f.x; x[q * 8 + e * 2 + 1] = f.y;
                }
            }
        } else {
#pragma unroll
            for (int e = 0; e < W; ++e) x[e] = (col + e < N) ? __bfloat162float(p[e]) : 0.f;
        }
    } else {
        const float* p = reinterpret_cast<const float*>(base) + (long long)row * ld + col;
        if (vec && col + W <= N) {
#pragma unroll
            for (int q = 0; q < W / 4; ++q) {
                float4 f = __ldg(reinterpret_cast<const float4*>(p) + q);
                x[q * 4] = f.x; x[q * 4 + 1] = f.y; x[q * 4 + 2] = f.z; x[q * 4 + 3] = f.w;
            }
        } else {
#pragma unroll
            for (int e = 0; e < W; ++e) x[e] = (col + e < N) ? p[e] : 0.f;
        }
    }
}

template <bool BF16, int W>
__device__ __forceinline__ void st_row(void* base, long long ld, int row, int col, int N, bool vec, const float (&x)[W]) {
    if constexpr (BF16) {
        __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(base) + (long long)row * ld + col;
        if (vec && col + W <= N) {
#pragma unroll
            for (int q = 0; q < W / 8; ++q) {
                uint4 u;
                __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(x[q * 8 + e * 2], x[q * 8 + e * 2 + 1]);
                reinterpret_cast<uint4*>(p)[q] = u;
            }
        } else {
#pragma unroll
            for (int e = 0; e < W; ++e) if (col + e < N) p[e] = __float2bfloat16_rn(x[e]);
        }
    } else {
        float* p = reinterpret_cast<float*>(base) + (long long)row * ld + col;
        if (vec && col + W <= N) {
#pragma unroll
            for (int q = 0; q < W / 4; ++q)
                reinterpret_cast<float4*>(p)[q] = make_float4(x[q * 4], x[q * 4 + 1], x[q * 4 + 2], x[q * 4 + 3]);
        } else {
#pragma unroll
            for (int e = 0; e < W; ++e) if (col + e < N) p[e] = x[e];
        }
    }
}

// ---------------------------------------------------------------------------------------------- fused epilogue
__device__ __forceinline__ bool epi_has_aux(const GemmParams& p) {
    return p.epi == EPI_BIAS_ACT_DZ || p.epi == EPI_BIAS_ACT_SE || p.epi == EPI_MUL_DACT || (p.epi == EPI_STORE && p.aux0 != nullptr);
}
__device__ __forceinline__ bool epi_has_out1(const GemmParams& p) { return p.epi == EPI_BIAS_ACT_DZ || p.epi == EPI_BIAS_ACT_SE; }

// Math for W consecutive columns [col, col+W) of one output row (every epilogue except EPI_ATOMIC).
//   in : v = accumulators, x = aux values (if epi_has_aux)        out: v = out0 values, x = out1 values (if epi_has_out1)
template <int W>
__device__ __forceinline__ void epi_math(const GemmParams& p, bool row_ok, int col, float (&v)[W], float (&x)[W], float& loss_acc) {
    switch (p.epi) {
        case EPI_STORE: {
#pragma unroll
            for (int e = 0; e < W; ++e) v[e] *= p.alpha;
            if (p.aux0 != nullptr) {
#pragma unroll
                for (int e = 0; e < W; ++e) v[e] = fmaf(p.beta, x[e], v[e]);
            }
        } break;
        case EPI_BIAS_ACT:
        case EPI_BIAS_ACT_DZ:
        case EPI_BIAS_ACT_SE: {
            if (p.bias != nullptr) {
                if (col + W <= p.N && (reinterpret_cast<uintptr_t>(p.bias + col) & 15) == 0) {   // warp-uniform broadcast loads
#pragma unroll
                    for (int g = 0; g < W / 4; ++g) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col) + g);
                        v[g * 4] += b4.x; v[g * 4 + 1] += b4.y; v[g * 4 + 2] += b4.z; v[g * 4 + 3] += b4.w;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < W; ++e) v[e] += (col + e < p.N) ? __ldg(p.bias + col + e) : 0.f;
                }
            }
#pragma unroll
            for (int e = 0; e < W; ++e) v[e] = act_apply(p.act, v[e]);
            if (p.epi == EPI_BIAS_ACT_DZ) {
#pragma unroll
                for (int e = 0; e < W; ++e) x[e] = x[e] * act_deriv_from_out(p.act, v[e]);
            } else if (p.epi == EPI_BIAS_ACT_SE) {
#pragma unroll
                for (int e = 0; e < W; ++e) {
                    const float d = (row_ok && col + e < p.N) ? x[e] - v[e] : 0.f;   // squaredError: NeuralNet.hs:61-68
                    loss_acc = fmaf(d, d, loss_acc);
                    x[e] = -2.0f * d * act_deriv_from_out(p.act, v[e]);
                }
            }
        } break;
        case EPI_MUL_DACT: {
#pragma unroll
            for (int e = 0; e < W; ++e) v[e] *= act_deriv_from_out(p.act, x[e]);
        } break;
        default: break;
    }
}

// Direct (register <-> global) epilogue: used for EPI_ATOMIC and whenever the outputs cannot be described by TMA
// tensor maps (unaligned base / leading dimension).  Row-per-thread accesses: correct everywhere, slow.
template <int W>
__device__ __forceinline__ void epi_direct(const GemmParams& p, int row, int col, bool vec, float (&v)[W], float& loss_acc) {
    const bool bf = p.io_bf16 != 0;
    if (p.epi == EPI_ATOMIC) {
        float* o = reinterpret_cast<float*>(p.out0) + (long long)row * p.ld_out0 + col;
        if (vec && col + W <= p.N) {
#pragma unroll
            for (int g = 0; g < W / 4; ++g)
                ptx::red_add_v4(o + g * 4, p.alpha * v[g * 4], p.alpha * v[g * 4 + 1], p.alpha * v[g * 4 + 2], p.alpha * v[g * 4 + 3]);
        } else {
#pragma unroll
            for (int e = 0; e < W; ++e) if (col + e < p.N) atomicAdd(o + e, p.alpha * v[e]);
        }
        return;
    }
    float x[W];
    if (epi_has_aux(p)) {
        if (bf) ld_row<true, W>(p.aux0, p.ld_aux0, row, col, p.N, vec, x);
        else ld_row<false, W>(p.aux0, p.ld_aux0, row, col, p.N, vec, x);
    }
    epi_math<W>(p, true, col, v, x, loss_acc);
    if (bf) st_row<true, W>(p.out0, p.ld_out0, row, col, p.N, vec, v);
    else st_row<false, W>(p.out0, p.ld_out0, row, col, p.N, vec, v);
    if (epi_has_out1(p)) {
        if (bf) st_row<true, W>(p.out1, p.ld_out1, row, col, p.N, vec, x);
        else st_row<false, W>(p.out1, p.ld_out1, row, col, p.N, vec, x);
    }
}

// Staging blocks are 32 rows x W elements of IO, rows of 64 or 128 bytes, laid out the way the TMA swizzle modes expect
// so that both the thread side (lane = row, 16-byte accesses) and the TMA side are bank-conflict free:
//   128-byte rows, SWIZZLE_128B: 16B chunk j of row r lives at r*128 + ((j ^ (r & 7)) << 4)
//    64-byte rows, SWIZZLE_64B : 16B chunk j of row r lives at r*64  + ((j ^ ((r >> 1) & 3)) << 4)
//    32-byte rows (thread-written only): chunk j of row r lives at r*32 + ((j ^ ((r >> 2) & 1)) << 4)
template <int ROWB> __device__ __forceinline__ int stage_off(int r, int j) {
    if constexpr (ROWB == 128) return r * 128 + ((j ^ (r & 7)) << 4);
    else if constexpr (ROWB == 64) return r * 64 + ((j ^ ((r >> 1) & 3)) << 4);
    else return r * 32 + ((j ^ ((r >> 2) & 1)) << 4);
}
template <typename IO, int W>
__device__ __forceinline__ void stage_read_row(uint32_t buf, int r, float (&x)[W]) {
    constexpr int ROWB = W * (int)sizeof(IO);
    static_assert(ROWB == 64 || ROWB == 128, "TMA-filled staging rows are 64 or 128 bytes");
#pragma unroll
    for (int j = 0; j < ROWB / 16; ++j) {
        const uint4 u = ptx::lds128(buf + stage_off<ROWB>(r, j));
        if constexpr (sizeof(IO) == 2) {
            const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(h[e]);
                x[j * 8 + e * 2] = f.x; x[j * 8 + e * 2 + 1] = f.y;
            }
        } else {
            x[j * 4] = __uint_as_float(u.x); x[j * 4 + 1] = __uint_as_float(u.y); x[j * 4 + 2] = __uint_as_float(u.z); x[j * 4 + 3] = __uint_as_float(u.w);
        }
    }
}
template <typename IO, int W>
__device__ __forceinline__ void stage_write_row(uint32_t buf, int r, const float (&x)[W]) {
    constexpr int ROWB = W * (int)sizeof(IO);
#pragma unroll
    for (int j = 0; j < ROWB / 16; ++j) {
        uint4 u;
        if constexpr (sizeof(IO) == 2) {
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(x[j * 8 + e * 2], x[j * 8 + e * 2 + 1]);
        } else {
            u = make_uint4(__float_as_uint(x[j * 4]), __float_as_uint(x[j * 4 + 1]), __float_as_uint(x[j * 4 + 2]), __float_as_uint(x[j * 4 + 3]));
        }
        ptx::sts128(buf + stage_off<ROWB>(r, j), u);
    }
}
// One row of W fp32 values -> an fp16 PAIR: t = x * s; hi = fp16(t), lo = fp16(t - hi)  (t - hi is exact in fp32).
// The two planes are staged as 32 x W fp16 blocks (64-byte rows, SWIZZLE_64B layout) at `buf` and `buf + 32 * W * 2`.
template <int W>
__device__ __forceinline__ void stage_write_row_f16pair(uint32_t buf, int r, const float (&x)[W], float s) {
    static_assert(W == 32 || W == 16, "64- or 32-byte fp16 rows");
#pragma unroll
    for (int j = 0; j < W / 8; ++j) {
        uint4 uh, ul;
        __half2* hh = reinterpret_cast<__half2*>(&uh);
        __half2* hl = reinterpret_cast<__half2*>(&ul);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float t0 = x[j * 8 + e * 2] * s, t1 = x[j * 8 + e * 2 + 1] * s;
            hh[e] = __floats2half2_rn(t0, t1);
            const float2 hf = __half22float2(hh[e]);
            hl[e] = __floats2half2_rn(t0 - hf.x, t1 - hf.y);
        }
        ptx::sts128(buf + stage_off<W * 2>(r, j), uh);
        ptx::sts128(buf + 32 * W * 2 + stage_off<W * 2>(r, j), ul);
    }
}
// Column sums of a staged 32 x W block, added into colsum[col .. col+W) with one red per column (db = sum_s dZ[s,:]).
template <typename IO, int W>
__device__ __forceinline__ void stage_colsum(uint32_t buf, int lane, int col, int N, float* colsum) {
    constexpr int ROWB = W * (int)sizeof(IO);
    constexpr int EPC = 16 / (int)sizeof(IO);          // elements per 16-byte chunk
    if constexpr (W == 16) {                           // fp32, 64-byte rows: lane = (column, row parity); the two parities meet by shuffle
        static_assert(sizeof(IO) == 4, "16-column blocks are fp32");
        const int c = lane & 15, h = lane >> 4;
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < 16; ++k)                   // row 2k + h: ((2k + h) >> 1) & 3 == k & 3, a compile-time constant per k
            part[k & 3] += ptx::lds32f(buf + (2 * k + h) * 64 + (((c >> 2) ^ (k & 3)) << 4) + (c & 3) * 4);
        float s = (part[0] + part[1]) + (part[2] + part[3]);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (h == 0 && col + c < N) atomicAdd(colsum + col + c, s);
        return;
    }
    static_assert(W == 32 || W == 16, "one column per lane");
    float part[4] = {0.f, 0.f, 0.f, 0.f};            // independent partial sums: no 32-long dependent FADD chain
    constexpr int PERIOD = 8;                          // stage_off(r + 8, j) == stage_off(r, j) + 8 * ROWB for both layouts
#pragma unroll
    for (int k = 0; k < PERIOD; ++k) {
        const uint32_t a = buf + stage_off<ROWB>(k, lane / EPC) + (lane % EPC) * (int)sizeof(IO);
#pragma unroll
        for (int m = 0; m < 32 / PERIOD; ++m) {
            if constexpr (sizeof(IO) == 2) part[m] += __uint_as_float(static_cast<uint32_t>(ptx::lds16(a + m * PERIOD * ROWB)) << 16);
            else part[m] += ptx::lds32f(a + m * PERIOD * ROWB);
        }
    }
    const float s = (part[0] + part[1]) + (part[2] + part[3]);
    if (col + lane < N) atomicAdd(colsum + col + lane, s);
}

// Coalesced copy of a staged 32 x W block to global memory: consecutive lanes take consecutive 16-byte chunks of a row, so a
// warp instruction writes whole 64/128-byte row segments (4-8 L1 wavefronts instead of the 32 of row-per-thread stores).
// Plain st.global: fire-and-forget, no completion to wait for before the staging block is reused.
template <typename IO, int W>
__device__ __forceinline__ void stage_store_global(uint32_t buf, int lane, void* base, long long ld, int row0, int col, int M, int N) {
    constexpr int ROWB = W * (int)sizeof(IO);
    constexpr int CPR = ROWB / 16;                 // 16-byte chunks per row
    constexpr int EPC = 16 / (int)sizeof(IO);      // elements per chunk
    constexpr int RPI = 32 / CPR;                  // rows covered by one warp instruction
    const int j = lane % CPR, c = col + j * EPC;
    const int rl = lane / CPR;
    IO* g0 = reinterpret_cast<IO*>(base) + (long long)(row0 + rl) * ld + c;
    if (row0 + 32 <= M && col + W <= N) {          // interior block (warp-uniform): no guards, two swizzle variants at most
        // rows rl + k*RPI.  128-byte rows: RPI = 4, the swizzle (r & 7) alternates between two values; 64-byte rows: RPI = 8, constant
        const uint32_t s0 = buf + stage_off<ROWB>(rl, j);
        const uint32_t s1 = buf + stage_off<ROWB>(rl + RPI, j);
        uint4 u[CPR];                              // all shared loads first, then all global stores: the LDS latency is paid once
#pragma unroll
        for (int k = 0; k < CPR; ++k) {
            const uint32_t a = (ROWB == 128) ? ((k & 1) ? s1 : s0) + (k >> 1) * (2 * RPI * ROWB) : s0 + k * (RPI * ROWB);
            u[k] = ptx::lds128(a);
        }
#pragma unroll
        for (int k = 0; k < CPR; ++k) *reinterpret_cast<uint4*>(g0 + (long long)k * RPI * ld) = u[k];
        return;
    }
    if (c >= N) return;
#pragma unroll
    for (int k = 0; k < CPR; ++k) {
        const int r = rl + k * RPI;
        if (row0 + r < M) {
            const uint4 u = ptx::lds128(buf + stage_off<ROWB>(r, j));
            IO* g = g0 + (long long)k * RPI * ld;
            if (c + EPC <= N) {
                *reinterpret_cast<uint4*>(g) = u;
            } else {
                const IO* e = reinterpret_cast<const IO*>(&u);
#pragma unroll
                for (int i = 0; i < EPC; ++i) if (c + i < N) g[i] = e[i];
            }
        }
    }
}

// Fused all-reduce of split-K partials: called by an epilogue warp after its local reds for one region (32 rows x `ncols` columns
// starting at (row0, col0)) of one work item.  The warp that completes the region last (all split_k partials are in) re-reads the
// finished sums from L2 and pushes them ONCE to the multicast alias: the switch adds them into every rank's gradient buffer.
__device__ __forceinline__ void mc_push_region_if_last(const GemmParams& p, int counter_idx, int lane, int row0, int col0, int ncols) {
    __threadfence();                                   // this lane's reds are performed before the arrival is counted
    __syncwarp();
    int last = 0;
    if (lane == 0) last = (atomicAdd(p.tile_counters + counter_idx, 1) == p.split_k - 1) ? 1 : 0;
    last = __shfl_sync(0xffffffffu, last, 0);
    if (!last) return;
    __threadfence();                                   // acquire: the other partials' reds (performed at L2) are visible to the loads below
    const float* src = reinterpret_cast<const float*>(p.out0);
    const bool vec = (reinterpret_cast<uintptr_t>(p.out0) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.out0_mc) & 15) == 0 && (p.ld_out0 & 3) == 0 && (col0 & 3) == 0;
    for (int r = 0; r < 32; ++r) {
        const int row = row0 + r;
        if (row >= p.M) break;
        const long long base = (long long)row * p.ld_out0 + col0;
        for (int c = lane * 4; c < ncols; c += 128) {
            if (vec && c + 4 <= ncols) {
                const float4 v = __ldcg(reinterpret_cast<const float4*>(src + base + c));
                ptx::multimem_red_add_v4(p.out0_mc + base + c, v.x, v.y, v.z, v.w);
            } else {
                for (int e = 0; e < 4 && c + e < ncols; ++e) ptx::multimem_red_add(p.out0_mc + base + c + e, __ldcg(src + base + c + e));
            }
        }
    }
}

// Per-warp state of the staged epilogue.  The aux operand (dA / target / previous activation / C) arrives by TMA into
// `aux_buf`; outputs are transposed through `out_buf` (== aux_buf in the 3-pass kernel, which has one block per warp).
struct EpiWarp {
    uint32_t aux_buf;      // shared-window byte addresses
    uint32_t out_buf;
    uint64_t* aux_bar;     // mbarrier the aux TMA load completes on
    uint32_t consumed;     // aux loads consumed so far (parity of the next wait)
    bool in_flight;        // the aux block of the block about to be processed has already been requested
    float out1_s;          // out1_pair: this lane's (row's) scale for the fp16 pair of out1
};

// The staging block was last READ by this warp with ld.shared (generic proxy) and is about to be WRITTEN by the TMA engine (async
// proxy): without a proxy fence the bulk copy can overtake those reads — measured: ~1e-3 of the dZ elements picked up the next
// sub-block's dA once the epilogue became fast enough (callers __syncwarp() after their reads, then every lane fences).
template <typename IO, int W>
__device__ __forceinline__ void epi_issue_aux(const CUtensorMap* tmAux, EpiWarp& w, int lane, int row0, int col) {
    ptx::fence_proxy_async_smem();
    if (lane == 0) {
        ptx::mbar_arrive_expect_tx(w.aux_bar, 32u * W * (uint32_t)sizeof(IO));
        ptx::tma_load_2d_s(w.aux_buf, tmAux, w.aux_bar, col, row0);
    }
    w.in_flight = true;
}

// One 32 x W block of the output tile, warp-collective: rows row0..row0+31 (lane = row), columns col..col+W-1.
//   v: this lane's accumulators.  next_col >= 0: column of the next block this warp will process in the same rows; when the
//   aux block is separate from the output block its aux operand is requested as soon as this block's has been read.
template <typename IO, int W, bool SHARED>
__device__ __forceinline__ void epi_block(const GemmParams& p, const CUtensorMap* tmAux, EpiWarp& w, int lane, int row0, int col, int next_col,
                                          float (&v)[W], float& loss_acc, volatile unsigned int* wd) {
    const bool has_aux = epi_has_aux(p), has_out1 = epi_has_out1(p);
    float x[W];
    if (has_aux) {
        if (!w.in_flight) epi_issue_aux<IO, W>(tmAux, w, lane, row0, col);
        ptx::mbar_wait(w.aux_bar, w.consumed & 1, wd, 0x600);
        ++w.consumed;
        stage_read_row<IO, W>(w.aux_buf, lane, x);
        __syncwarp();
        w.in_flight = false;
        if (!SHARED && next_col >= 0) epi_issue_aux<IO, W>(tmAux, w, lane, row0, next_col);
    }
    epi_math<W>(p, row0 + lane < p.M, col, v, x, loss_acc);
    stage_write_row<IO, W>(w.out_buf, lane, v);
    __syncwarp();
    stage_store_global<IO, W>(w.out_buf, lane, p.out0, p.ld_out0, row0, col, p.M, p.N);
    if (p.colsum != nullptr && p.colsum_src == 1) stage_colsum<IO, W>(w.out_buf, lane, col, p.N, p.colsum);
    __syncwarp();
    if (has_out1) {
        if constexpr (std::is_same<IO, float>::value) {
            if (p.out1_pair) {   // out1 leaves as an fp16 pair (operand of the next F16X3 GEMMs); its column sums come from the fp32 values
                if (p.colsum != nullptr && p.colsum_src == 2) {
                    stage_write_row<IO, W>(w.out_buf, lane, x);
                    __syncwarp();
                    stage_colsum<IO, W>(w.out_buf, lane, col, p.N, p.colsum);
                    __syncwarp();
                }
                stage_write_row_f16pair<W>(w.out_buf, lane, x, w.out1_s);
                __syncwarp();
                stage_store_global<__half, W>(w.out_buf, lane, p.out1, p.ld_out1, row0, col, p.M, p.N);
                stage_store_global<__half, W>(w.out_buf + 32 * W * 2, lane, p.out1b, p.ld_out1, row0, col, p.M, p.N);
                __syncwarp();
                return;
            }
        }
        stage_write_row<IO, W>(w.out_buf, lane, x);
        __syncwarp();
        stage_store_global<IO, W>(w.out_buf, lane, p.out1, p.ld_out1, row0, col, p.M, p.N);
        if (p.colsum != nullptr && p.colsum_src == 2) stage_colsum<IO, W>(w.out_buf, lane, col, p.N, p.colsum);
        __syncwarp();
    }
}

// Interior blocks only (all 32 rows and W columns valid): staged 32 x W block -> global, no guards.  Valid for 64- and 32-byte
// staging rows, whose swizzle term is the same for all rows one warp instruction apart (see stage_off).
template <typename IO, int W>
__device__ __forceinline__ void stage_store_interior(uint32_t buf, int lane, void* base, long long ld, int row0, int col) {
    constexpr int ROWB = W * (int)sizeof(IO), CPR = ROWB / 16, EPC = 16 / (int)sizeof(IO), RPI = 32 / CPR;
    static_assert(ROWB == 64 || ROWB == 32, "constant-swizzle layouts");
    const int j = lane % CPR, rl = lane / CPR;
    IO* g0 = reinterpret_cast<IO*>(base) + (long long)(row0 + rl) * ld + col + j * EPC;
    const uint32_t s0 = buf + stage_off<ROWB>(rl, j);
    uint4 u[CPR];
#pragma unroll
    for (int k = 0; k < CPR; ++k) u[k] = ptx::lds128(s0 + k * (RPI * ROWB));
#pragma unroll
    for (int k = 0; k < CPR; ++k) *reinterpret_cast<uint4*>(g0 + (long long)k * RPI * ld) = u[k];
}

// The hot epilogue of the F16X3 forward GEMM, written out without any of the generality of epi_block: interior 32 x 16 sub-blocks of
//   A = logistic(acc + b),  dZ = dA * A (1 - A),  db += column sums of dZ,  (dZ1, dZ2) = fp16 pair of dZ * s.
// (The generic path costs ~680 instructions per sub-block — every epilogue variant, dtype and edge case is compiled into one
// stream that no longer fits the instruction cache; this one is ~410.)  v: in = scaled accumulators + bias, lane = row.
// It works on a PAIR of sub-blocks (32 columns) at a time: both staging blocks of the warp serve both sub-blocks in every
// phase (dA in, A out, dZ for the column sums, the fp16 pair out), so a 32-column step has 7 warp-synchronised phases instead of
// 14 and every phase has twice the independent work in flight (the 8 epilogue warps are latency-bound: 2 per scheduler).
// The two dA blocks of the NEXT pair are requested (one mbarrier, 4 KiB) when this pair's last staged bytes have been read, and
// their latency hides behind the logistic of that pair, which does not need them.
__device__ __forceinline__ void epi_issue_aux_pair(const CUtensorMap* tmAux, EpiWarp& w, int lane, int row0, int col) {
    ptx::fence_proxy_async_smem();
    if (lane == 0) {
        ptx::mbar_arrive_expect_tx(w.aux_bar, 2u * 32u * 16u * 4u);
        ptx::tma_load_2d_s(w.aux_buf, tmAux, w.aux_bar, col, row0);
        ptx::tma_load_2d_s(w.out_buf, tmAux, w.aux_bar, col + 16, row0);
    }
    w.in_flight = true;
}
__device__ __forceinline__ void epi_fwd_pair_lean2(const GemmParams& p, const CUtensorMap* tmAux, EpiWarp& w, int lane, int row0, int col, int next_col,
                                                   float (&v)[32], volatile unsigned int* wd) {
    if (!w.in_flight) epi_issue_aux_pair(tmAux, w, lane, row0, col);
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = act_apply(ACT_LOGISTIC, v[e]);       // A: needs no dA — covers the TMA latency
    ptx::mbar_wait(w.aux_bar, w.consumed & 1, wd, 0x600);
    ++w.consumed;
    w.in_flight = false;
    float x[32];
    stage_read_row<float, 16>(w.aux_buf, lane, *reinterpret_cast<float (*)[16]>(&x[0]));
    stage_read_row<float, 16>(w.out_buf, lane, *reinterpret_cast<float (*)[16]>(&x[16]));
    __syncwarp();
#pragma unroll
    for (int e = 0; e < 32; ++e) x[e] = x[e] * (v[e] * (1.0f - v[e]));
    // ---- A out
    stage_write_row<float, 16>(w.aux_buf, lane, *reinterpret_cast<float (*)[16]>(&v[0]));
    stage_write_row<float, 16>(w.out_buf, lane, *reinterpret_cast<float (*)[16]>(&v[16]));
    __syncwarp();
    stage_store_interior<float, 16>(w.aux_buf, lane, p.out0, p.ld_out0, row0, col);
    stage_store_interior<float, 16>(w.out_buf, lane, p.out0, p.ld_out0, row0, col + 16);
    __syncwarp();
    // ---- db: lanes 0-15 sum the 16 columns of the first block, lanes 16-31 those of the second (rows visited in opposite order of
    //      the row pair so that the two half-warps hit different banks)
    stage_write_row<float, 16>(w.aux_buf, lane, *reinterpret_cast<float (*)[16]>(&x[0]));
    stage_write_row<float, 16>(w.out_buf, lane, *reinterpret_cast<float (*)[16]>(&x[16]));
    __syncwarp();
    {
        const int c = lane & 15, h = lane >> 4;
        const uint32_t buf = h ? w.out_buf : w.aux_buf;
        float part[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const int rr = r ^ h;                      // ((rr >> 1) & 3) == ((r >> 1) & 3): a compile-time swizzle term
            part[r & 3] += ptx::lds32f(buf + rr * 64 + (((c >> 2) ^ ((r >> 1) & 3)) << 4) + (c & 3) * 4);
        }
        atomicAdd(p.colsum + col + lane, (part[0] + part[1]) + (part[2] + part[3]));
    }
    __syncwarp();
    // ---- fp16 pair out
    stage_write_row_f16pair<16>(w.aux_buf, lane, *reinterpret_cast<float (*)[16]>(&x[0]), w.out1_s);
    stage_write_row_f16pair<16>(w.out_buf, lane, *reinterpret_cast<float (*)[16]>(&x[16]), w.out1_s);
    __syncwarp();
    stage_store_interior<__half, 16>(w.aux_buf, lane, p.out1, p.ld_out1, row0, col);
    stage_store_interior<__half, 16>(w.aux_buf + 32 * 16 * 2, lane, p.out1b, p.ld_out1, row0, col);
    stage_store_interior<__half, 16>(w.out_buf, lane, p.out1, p.ld_out1, row0, col + 16);
    stage_store_interior<__half, 16>(w.out_buf + 32 * 16 * 2, lane, p.out1b, p.ld_out1, row0, col + 16);
    __syncwarp();
    if (next_col >= 0) epi_issue_aux_pair(tmAux, w, lane, row0, next_col);
}

// 8 fp32 values -> 8 x bf16(x) and 8 x bf16(x - trunc_tf32(x))   (operands of the two bf16 correction passes)
__device__ __forceinline__ void split8_bf16(const uint4& x0, const uint4& x1, uint4& v16, uint4& l16) {
    const float f[8] = {__uint_as_float(x0.x), __uint_as_float(x0.y), __uint_as_float(x0.z), __uint_as_float(x0.w),
                        __uint_as_float(x1.x), __uint_as_float(x1.y), __uint_as_float(x1.z), __uint_as_float(x1.w)};
    __nv_bfloat162* hv = reinterpret_cast<__nv_bfloat162*>(&v16);
    __nv_bfloat162* hl = reinterpret_cast<__nv_bfloat162*>(&l16);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float a0 = f[2 * e], a1 = f[2 * e + 1];
        hv[e] = __floats2bfloat162_rn(a0, a1);
        hl[e] = __floats2bfloat162_rn(a0 - __uint_as_float(__float_as_uint(a0) & 0xffffe000u), a1 - __uint_as_float(__float_as_uint(a1) & 0xffffe000u));
    }
}

__device__ __forceinline__ void split4_bf16(const uint4& x, uint2& v8, uint2& l8) {
    const float f[4] = {__uint_as_float(x.x), __uint_as_float(x.y), __uint_as_float(x.z), __uint_as_float(x.w)};
    __nv_bfloat162* hv = reinterpret_cast<__nv_bfloat162*>(&v8);
    __nv_bfloat162* hl = reinterpret_cast<__nv_bfloat162*>(&l8);
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const float a0 = f[2 * e], a1 = f[2 * e + 1];
        hv[e] = __floats2bfloat162_rn(a0, a1);
        hl[e] = __floats2bfloat162_rn(a0 - __uint_as_float(__float_as_uint(a0) & 0xffffe000u), a1 - __uint_as_float(__float_as_uint(a1) & 0xffffe000u));
    }
}

// EV = 1 (fp16-pair kernels, A K-major): the epilogue variant of the MLP's hidden layers — specialised interior-tile code for
//   EPI_BIAS_ACT / logistic (B K-major: the forward GEMM) and EPI_MUL_DACT / logistic (B MN-major: the dA GEMM) instead of the
//   ffLayer step's paired forward epilogue and plain store (EV = 0).  Separate instantiations: the hot ffLayer kernels keep their
//   register allocation and instruction footprint (both were measured to be sensitive to any code added next to them).
template <typename T, int MA, int MB, int BN, int STAGES, int PASSES, int CG, int EV = 0>
__global__ void __launch_bounds__((GemmCfg<T, MA, MB, BN, STAGES, PASSES, CG>::NUM_THREADS), 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmAux,
                 const __grid_constant__ CUtensorMap tmB16, const __grid_constant__ CUtensorMap tmBlo16, const GemmParams p) {
    using Cfg = GemmCfg<T, MA, MB, BN, STAGES, PASSES, CG>;
    constexpr bool kBF16 = sizeof(T) == 2;   // 16-bit operands (bf16 or fp16): kind::f16, K = 16 per instruction
    constexpr uint32_t kFmt = std::is_same<T, __half>::value ? 0u : (kBF16 ? 1u : 2u);   // instruction-descriptor operand format
    constexpr bool kPresplit = PASSES == 4;  // F16X3: (hi, lo) fp16 planes of both operands arrive by TMA (maps tmB16 / tmBlo16 = A_lo / B_lo)
    constexpr int NPASS = Cfg::NPASS;
    constexpr bool kSplitter = PASSES == 2 || PASSES == 3;
    constexpr bool kChunked = PASSES >= 2;   // TMEM holds one chunk; the running sum lives in epilogue registers
    constexpr int BM = Cfg::BM;
    constexpr int KB = Cfg::KB_ELEMS;
    constexpr int kSplitWarp0 = 12;          // first splitter warp (3-pass layout)

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;                       // [STAGES]  TMA landed
    uint64_t* empty_bar = bars + STAGES;             // [STAGES]  MMAs reading the stage retired
    uint64_t* ready_bar = bars + 2 * STAGES;         // [STAGES]  hi/lo split done (3-pass)
    uint64_t* tmem_full = bars + 3 * STAGES;         // [2]
    uint64_t* tmem_empty = bars + 3 * STAGES + 2;    // [2]
    uint64_t* epi_bar = bars + 3 * STAGES + 4;       // [EPI_WARPS] aux-operand TMA loads of the epilogue warps
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 3 * STAGES + 4 + Cfg::EPI_WARPS);
    uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES + 1024;   // barriers occupy the first BAR_BYTES of a 1 KiB slot so the staging blocks stay 1024-aligned

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    volatile unsigned int* wd = p.watchdog;
    // CTA pair (CG = 2): the two CTAs of a cluster own rows [0,128) and [128,256) of a 256-row tile and half of the B tile each;
    // the leader (rank 0) issues tcgen05.mma.cta_group::2 for both.  All cross-CTA signalling goes through mbarriers of the leader
    // (remote arrives) or through multicast commits (same barrier offset in both CTAs).
    const uint32_t cta_rank = CG == 2 ? ptx::cluster_ctarank() : 0u;
    const bool leader = cta_rank == 0;
    const int wid0 = blockIdx.x / CG, wstride = gridDim.x / CG;   // work items are distributed over CTA groups

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&tmA);
        ptx::prefetch_tensormap(&tmB);
        if (p.tma_epi && epi_has_aux(p)) ptx::prefetch_tensormap(&tmAux);
        if ((PASSES == 2 && p.b_presplit) || kPresplit) { ptx::prefetch_tensormap(&tmB16); ptx::prefetch_tensormap(&tmBlo16); }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&empty_bar[s], 1);
            // ready: CG = 1: the 128 splitter threads (3-pass).  CG = 2: splitter threads of both CTAs (3-pass) or the two
            // relay threads that forward "my TMA data has landed" to the leader (1-pass)
            ptx::mbar_init(&ready_bar[s], CG == 1 ? 128 : (kSplitter ? 256 : 2));
        }
        for (int a = 0; a < 2; ++a) {
            ptx::mbar_init(&tmem_full[a], 1);
            ptx::mbar_init(&tmem_empty[a], CG * Cfg::EPI_WARPS);   // one elected lane per epilogue warp (of both CTAs)
        }
        for (int e = 0; e < Cfg::EPI_WARPS; ++e) ptx::mbar_init(&epi_bar[e], 1);
        ptx::fence_barrier_init();
    }
    if (warp == 2) {
        if constexpr (CG == 2) { ptx::tmem_alloc_2cta(tmem_ptr, Cfg::TMEM_COLS); ptx::tmem_relinquish_2cta(); }
        else { ptx::tmem_alloc(tmem_ptr, Cfg::TMEM_COLS); ptx::tmem_relinquish(); }
    }
    ptx::tcgen05_fence_before();
    __syncthreads();
    if constexpr (CG == 2) ptx::cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
    ptx::tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const int total_work = num_tiles * p.split_k;
    const int chunk_kb = kChunked ? p.chunk_kb : (1 << 30);
    const int chunk_head_kb = kChunked ? max(p.chunk_head_kb, p.chunk_kb) : (1 << 28);
    // chunk boundaries of a work item [kb0, kb1): two head chunks, then chunk_kb each — the MMA warp and the epilogue warps agree on it
    auto chunk_len = [&](int kc, int kb0) { return (kc - kb0 < 2 * chunk_head_kb) ? chunk_head_kb : chunk_kb; };

    if (warp < 4) {
        if constexpr (kChunked) ptx::setmaxnreg_dec<48>();
        else ptx::setmaxnreg_dec<56>();
        if (warp == 0 && lane == 0) {
            // ===================================================== TMA producer
            int s = 0; uint32_t ph = 0;
            for (int w = wid0; w < total_work; w += wstride) {
                const int tile = w % num_tiles, split = w / num_tiles;
                const int m0 = (tile / p.num_n_tiles) * (BM * CG) + (int)cta_rank * BM;            // this CTA's 128 rows
                const int n0 = (tile % p.num_n_tiles) * BN + (int)cta_rank * (BN / CG);            // this CTA's share of the B tile
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    if constexpr (CG == 2) ptx::mbar_wait_cluster(&empty_bar[s], ph ^ 1, wd, 0x100 + s);
                    else ptx::mbar_wait(&empty_bar[s], ph ^ 1, wd, 0x100 + s);
                    ptx::mbar_arrive_expect_tx(&full_bar[s], kPresplit ? 2 * Cfg::RAW_BYTES : Cfg::RAW_BYTES + ((PASSES == 2 && p.b_presplit) ? Cfg::B_BYTES : 0));
                    uint8_t* sa = smem + s * Cfg::STAGE_BYTES;
                    uint8_t* sb = sa + Cfg::A_BYTES;
                    auto load_a = [&](uint8_t* dst, const CUtensorMap* map) {
                        if constexpr (MA == MAJOR_K) {
                            ptx::tma_load_2d(dst, map, &full_bar[s], kb * KB, m0);
                        } else {
#pragma unroll
                            for (int r = 0; r < BM / KB; ++r)
                                ptx::tma_load_2d(dst + r * (KB * 128), map, &full_bar[s], m0 + r * KB, kb * KB);
                        }
                    };
                    auto load_b = [&](uint8_t* dst, const CUtensorMap* map) {
                        if constexpr (MB == MAJOR_K) {
                            ptx::tma_load_2d(dst, map, &full_bar[s], kb * KB, n0);
                        } else {
#pragma unroll
                            for (int r = 0; r < (BN / CG) / KB; ++r)
                                ptx::tma_load_2d(dst + r * (KB * 128), map, &full_bar[s], n0 + r * KB, kb * KB);
                        }
                    };
                    load_a(sa, &tmA);
                    load_b(sb, &tmB);
                    if constexpr (kPresplit) {   // the lo planes land where the 3-pass splitter would have written its lo tiles
                        load_a(sa + Cfg::RAW_BYTES, &tmB16);
                        load_b(sb + Cfg::RAW_BYTES, &tmBlo16);
                    }
                    if constexpr (PASSES == 2) {
                        if (p.b_presplit) {   // the small operand was split once in HBM: its bf16 tiles arrive ready-made
                            uint8_t* b16 = sa + Cfg::RAW_BYTES + Cfg::A_BYTES;
                            uint8_t* blo16 = b16 + Cfg::B_BYTES / 2;
                            if constexpr (MB == MAJOR_K) {
                                ptx::tma_load_2d(b16, &tmB16, &full_bar[s], kb * KB, n0);
                                ptx::tma_load_2d(blo16, &tmBlo16, &full_bar[s], kb * KB, n0);
                            } else {
#pragma unroll
                                for (int r = 0; r < (BN / CG) / 64; ++r) {
                                    ptx::tma_load_2d(b16 + r * 4096, &tmB16, &full_bar[s], n0 + r * 64, kb * KB);
                                    ptx::tma_load_2d(blo16 + r * 4096, &tmBlo16, &full_bar[s], n0 + r * 64, kb * KB);
                                }
                            }
                        }
                    }
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        } else if (warp == 1 && lane == 0 && leader) {
            // ===================================================== MMA issuer (CG = 2: the leader CTA issues for the pair)
            constexpr uint32_t idesc = ptx::make_idesc(kFmt, MA == MAJOR_MN, MB == MAJOR_MN, BM * CG, BN);
            // byte advance of the descriptor start address per UMMA_K step, and the LBO/SBO of each layout
            constexpr uint32_t a_step = (MA == MAJOR_K) ? 32u : (uint32_t)Cfg::UMMA_K * 128u;
            constexpr uint32_t b_step = (MB == MAJOR_K) ? 32u : (uint32_t)Cfg::UMMA_K * 128u;
            constexpr uint32_t a_lbo = (MA == MAJOR_K) ? 16u : (uint32_t)KB * 128u;
            constexpr uint32_t b_lbo = (MB == MAJOR_K) ? 16u : (uint32_t)KB * 128u;
            // MN-major 32-bit operands must use the 32B-atom flavour of the 128B swizzle (4-row atoms, SBO = 512)
            constexpr uint32_t a_lt = (MA == MAJOR_MN && !kBF16) ? 1u : 2u, b_lt = (MB == MAJOR_MN && !kBF16) ? 1u : 2u;
            constexpr uint32_t a_sbo = a_lt == 1u ? 512u : 1024u, b_sbo = b_lt == 1u ? 512u : 1024u;
            int s = 0; uint32_t ph = 0; int it = 0;   // `it` counts accumulation units: work items (1-pass) or chunks (3-pass)
            for (int w = wid0; w < total_work; w += wstride) {
                const int split = w / num_tiles;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
                for (int kc = kb0, kce; kc < kb1; kc = kce, ++it) {
                    kce = min(kc + chunk_len(kc, kb0), kb1);
                    const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
                    if constexpr (CG == 2) ptx::mbar_wait_cluster(&tmem_empty[acc], acc_ph ^ 1, wd, 0x200 + acc);
                    else ptx::mbar_wait(&tmem_empty[acc], acc_ph ^ 1, wd, 0x200 + acc);
                    ptx::tcgen05_fence_after();
                    const uint32_t d_tmem = tmem_base + acc * BN;
                    uint32_t first = 1;
                    for (int kb = kc; kb < kce; ++kb) {
                        // CG = 2: `ready` collects both CTAs (relay threads / splitter threads); the leader's own `full` is implied
                        if constexpr (CG == 2) ptx::mbar_wait_cluster(&ready_bar[s], ph, wd, 0x300 + s);
                        else ptx::mbar_wait(&full_bar[s], ph, wd, 0x300 + s);
                        ptx::tcgen05_fence_after();
                        const uint32_t sa = ptx::smem_u32(smem + s * Cfg::STAGE_BYTES);
                        const uint32_t sb = sa + Cfg::A_BYTES;
                        if constexpr (PASSES == 2) {
                            // hi*hi in TF32 from the raw tiles (the hardware truncates), then the two first-order corrections
                            //   bf16(A_lo) * bf16(B)   and   bf16(A) * bf16(B_lo)
                            // as bf16 MMAs: K = 16 per instruction, so each correction pass costs half a TF32 pass.  Their rounding
                            // error (2^-9 on a term that is 2^-11 of the product) is ~2^-20 per product, unbiased.
                            if constexpr (CG == 1) { ptx::mbar_wait(&ready_bar[s], ph, wd, 0x340 + s); ptx::tcgen05_fence_after(); }
                            constexpr uint32_t idesc16 = ptx::make_idesc(1u, MA == MAJOR_MN, MB == MAJOR_MN, BM * CG, BN);
                            const uint32_t a16 = sa + Cfg::RAW_BYTES, alo16 = a16 + Cfg::A_BYTES / 2;
                            const uint32_t b16 = alo16 + Cfg::A_BYTES / 2, blo16 = b16 + Cfg::B_BYTES / 2;
#pragma unroll
                            for (int j = 0; j < Cfg::KSTEPS; ++j) {
                                const uint64_t ad = ptx::make_smem_desc_sw128(sa + j * a_step, a_lbo, a_sbo, a_lt);
                                const uint64_t bd = ptx::make_smem_desc_sw128(sb + j * b_step, b_lbo, b_sbo, b_lt);
                                if constexpr (CG == 2) ptx::umma_tf32_2cta(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                else ptx::umma_tf32(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                first = 0;
                            }
                            // bf16 tiles written by the splitter.  K-major: rows = MN index, 64 bytes each, SWIZZLE_64B (8-row groups of 512 B),
                            // 32 bytes per K = 16 step.  MN-major: 64-element MN groups of [32 k-rows x 128 B], SWIZZLE_128B (LBO = 4096 between
                            // groups, SBO = 1024 between 8-row atoms), 16 k-rows = 2048 bytes per step.
                            constexpr uint32_t a16_lt = MA == MAJOR_K ? 4u : 2u, b16_lt = MB == MAJOR_K ? 4u : 2u;
                            constexpr uint32_t a16_lbo = MA == MAJOR_K ? 16u : 4096u, b16_lbo = MB == MAJOR_K ? 16u : 4096u;
                            constexpr uint32_t a16_sbo = MA == MAJOR_K ? 512u : 1024u, b16_sbo = MB == MAJOR_K ? 512u : 1024u;
                            constexpr uint32_t a16_step = MA == MAJOR_K ? 32u : 2048u, b16_step = MB == MAJOR_K ? 32u : 2048u;
#pragma unroll
                            for (int pass = 1; pass <= 2; ++pass) {
                                const uint32_t pa = pass == 1 ? alo16 : a16, pb = pass == 1 ? b16 : blo16;
#pragma unroll
                                for (int j = 0; j < 2; ++j) {   // 32 K-elements = 2 x (K = 16)
                                    const uint64_t ad = ptx::make_smem_desc_sw128(pa + j * a16_step, a16_lbo, a16_sbo, a16_lt);
                                    const uint64_t bd = ptx::make_smem_desc_sw128(pb + j * b16_step, b16_lbo, b16_sbo, b16_lt);
                                    if constexpr (CG == 2) ptx::umma_f16_2cta(d_tmem, ad, bd, idesc16, 1u);
                                    else ptx::umma_f16(d_tmem, ad, bd, idesc16, 1u);
                                }
                            }
                        } else {
#pragma unroll
                        for (int pass = 0; pass < NPASS; ++pass) {
                            // pass 0: A_hi*B_hi   pass 1: A_lo*B_hi   pass 2: A_hi*B_lo
                            // pass 0 needs only the raw tiles, so it is issued as soon as the TMA data lands and runs on the
                            // tensor pipe while the splitter warps are still producing the lo tiles for passes 1 and 2
                            if (PASSES == 3 && pass == 1 && CG == 1) {
                                ptx::mbar_wait(&ready_bar[s], ph, wd, 0x340 + s);
                                ptx::tcgen05_fence_after();
                            }
                            const uint32_t pa = sa + (pass == 1 ? Cfg::RAW_BYTES : 0);
                            const uint32_t pb = sb + (pass == 2 ? Cfg::RAW_BYTES : 0);
#pragma unroll
                            for (int j = 0; j < Cfg::KSTEPS; ++j) {
                                const uint64_t ad = ptx::make_smem_desc_sw128(pa + j * a_step, a_lbo, a_sbo, a_lt);
                                const uint64_t bd = ptx::make_smem_desc_sw128(pb + j * b_step, b_lbo, b_sbo, b_lt);
                                if constexpr (CG == 2) {
                                    if constexpr (kBF16) ptx::umma_f16_2cta(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                    else ptx::umma_tf32_2cta(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                } else {
                                    if constexpr (kBF16) ptx::umma_f16(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                    else ptx::umma_tf32(d_tmem, ad, bd, idesc, first ? 0u : 1u);
                                }
                                first = 0;
                            }
                        }
                        }
                        if constexpr (CG == 2) ptx::umma_commit_2cta(&empty_bar[s], 0x3);   // frees the stage in both CTAs
                        else ptx::umma_commit(&empty_bar[s]);
                        if (++s == STAGES) { s = 0; ph ^= 1; }
                    }
                    if constexpr (CG == 2) ptx::umma_commit_2cta(&tmem_full[acc], 0x3);
                    else ptx::umma_commit(&tmem_full[acc]);
                }
            }
        } else if (CG == 2 && !kSplitter && warp == 3 && lane == 0) {
            // ===================================================== relay (CTA pair, 1-pass): forward "my stage has landed" to the leader
            const uint32_t ready0 = ptx::mapa(ptx::smem_u32(&ready_bar[0]), 0);
            int s = 0; uint32_t ph = 0;
            for (int w = wid0; w < total_work; w += wstride) {
                const int split = w / num_tiles;
                const int kb0 = split * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
                for (int kb = kb0; kb < kb1; ++kb) {
                    ptx::mbar_wait(&full_bar[s], ph, wd, 0x700 + s);
                    ptx::mbar_arrive_cluster(ready0 + s * 8);
                    if (++s == STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if constexpr (!kChunked) {
        // ===================================================== epilogue, 1-pass: warps 4-11, one TMEM buffer per work item
        ptx::setmaxnreg_inc<224>();
        constexpr int W = Cfg::EPI_W;
        constexpr int HC = BN / 2;            // columns per warp: half of the tile
        const int q = warp & 3;               // TMEM lane quarter this warp may read (hardware rule: lanes 32*(warp%4)..+31)
        const int half = (warp - 4) >> 2;     // column half
        const bool vec = p.vec_ok != 0;
        const bool tma = p.tma_epi != 0;
        EpiWarp ew;
        ew.aux_buf = ptx::smem_u32(epi_smem + (warp - 4) * Cfg::EPI_WARP_BYTES);
        ew.out_buf = ew.aux_buf + (Cfg::EPI_NBUF - 1) * Cfg::EPI_BLOCK_BYTES;
        ew.aux_bar = &epi_bar[warp - 4]; ew.consumed = 0; ew.in_flight = false; ew.out1_s = 1.0f;
        const uint32_t tmem_empty0 = CG == 2 ? ptx::mapa(ptx::smem_u32(&tmem_empty[0]), 0) : 0u;   // the leader's accumulator-free barriers
        int it = 0;
        float loss_acc = 0.f;
        for (int w = wid0; w < total_work; w += wstride, ++it) {
            const int tile = w % num_tiles;
            const int m0 = (tile / p.num_n_tiles) * (BM * CG) + (int)cta_rank * BM, n0 = (tile % p.num_n_tiles) * BN + half * HC;
            const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
            const int row0 = m0 + q * 32;
            const int row = row0 + lane;
            const int ncols = min(HC, p.N - n0);                              // valid columns of this warp's half (may be <= 0)
            const int nblk = (row0 < p.M && ncols > 0) ? (ncols + W - 1) / W : 0;   // W-column blocks this warp owns in this tile
            if (tma && nblk > 0 && epi_has_aux(p)) {                          // the aux operand does not depend on the accumulators:
                epi_issue_aux<T, W>(&tmAux, ew, lane, row0, n0);              // first block straight into the staging block,
                if (lane > 0 && lane < nblk) ptx::tma_prefetch_l2_2d(&tmAux, n0 + lane * W, row0);   // the others into L2 meanwhile
            }
            if constexpr (CG == 2) ptx::mbar_wait_cluster(&tmem_full[acc], acc_ph, wd, 0x400 + acc);
            else ptx::mbar_wait(&tmem_full[acc], acc_ph, wd, 0x400 + acc);
            ptx::tcgen05_fence_after();
            const uint32_t t_row = tmem_base + acc * BN + half * HC + (static_cast<uint32_t>(q * 32) << 16);
            // single-pass (TF32 / bf16) forward GEMM of the ffLayer step on an interior tile: the specialised block code — same steps as
            // epi_block without its generality (every epilogue variant, dtype and edge case in one stream), the TMEM load of block
            // c+1 in flight while block c is processed, and the accumulator buffer handed back as soon as the last load has landed
            bool lean1 = false;
            if constexpr (!std::is_same<T, __half>::value && !Cfg::EPI_SHARED) {   // fp32 (TF32) and bf16 operands: staged blocks have the operand dtype
                lean1 = tma && p.epi == EPI_BIAS_ACT_DZ && p.act == ACT_LOGISTIC && !p.out1_pair && p.colsum != nullptr && p.colsum_src == 2 &&
                        p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 && row0 + 32 <= p.M && n0 + HC <= p.N && !(p.debug & 1);
                if (lean1) {
                    uint32_t raw[2][32];
                    ptx::tmem_ld_32x32b_x32(t_row, raw[0]);
#pragma unroll
                    for (int c = 0; c < HC / 32; ++c) {
                        ptx::tmem_ld_wait();
                        if (c + 1 < HC / 32) ptx::tmem_ld_32x32b_x32(t_row + (c + 1) * 32, raw[(c + 1) & 1]);
                        else {                                  // every column of the accumulator is in registers: free the buffer now
                            ptx::tcgen05_fence_before();
                            __syncwarp();
                            if (lane == 0) { if constexpr (CG == 2) ptx::mbar_arrive_cluster(tmem_empty0 + acc * 8); else ptx::mbar_arrive(&tmem_empty[acc]); }
                        }
                        const int col = n0 + c * 32;
                        float v[32], x[32];
                        const float4* b4p = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
                        for (int g = 0; g < 8; ++g) {
                            const float4 b4 = __ldg(b4p + g);
                            v[g * 4] = act_apply(ACT_LOGISTIC, __uint_as_float(raw[c & 1][g * 4]) + b4.x);
                            v[g * 4 + 1] = act_apply(ACT_LOGISTIC, __uint_as_float(raw[c & 1][g * 4 + 1]) + b4.y);
                            v[g * 4 + 2] = act_apply(ACT_LOGISTIC, __uint_as_float(raw[c & 1][g * 4 + 2]) + b4.z);
                            v[g * 4 + 3] = act_apply(ACT_LOGISTIC, __uint_as_float(raw[c & 1][g * 4 + 3]) + b4.w);
                        }
                        if (!ew.in_flight) epi_issue_aux<T, 32>(&tmAux, ew, lane, row0, col);
                        ptx::mbar_wait(ew.aux_bar, ew.consumed & 1, wd, 0x600);
                        ++ew.consumed;
                        stage_read_row<T, 32>(ew.aux_buf, lane, x);
                        __syncwarp();
                        ew.in_flight = false;
                        if (c + 1 < HC / 32) epi_issue_aux<T, 32>(&tmAux, ew, lane, row0, col + 32);
#pragma unroll
                        for (int e = 0; e < 32; ++e) x[e] = x[e] * (v[e] * (1.0f - v[e]));
                        stage_write_row<T, 32>(ew.out_buf, lane, v);
                        __syncwarp();
                        stage_store_global<T, 32>(ew.out_buf, lane, p.out0, p.ld_out0, row0, col, p.M, p.N);
                        __syncwarp();
                        stage_write_row<T, 32>(ew.out_buf, lane, x);
                        __syncwarp();
                        stage_colsum<T, 32>(ew.out_buf, lane, col, p.N, p.colsum);
                        stage_store_global<T, 32>(ew.out_buf, lane, p.out1, p.ld_out1, row0, col, p.M, p.N);
                        __syncwarp();
                    }
                    continue;                                   // next work item (the accumulator was released above)
                }
            }
#pragma unroll 1
            for (int c = 0; c < nblk; ++c) {
                uint32_t raw[W];
                ptx::tmem_ld_32x32b_x32(t_row + c * W, raw);
                ptx::tmem_ld_wait();
                const int col = n0 + c * W;
                float v[W];
#pragma unroll
                for (int e = 0; e < W; ++e) v[e] = __uint_as_float(raw[e]);
                if (tma) epi_block<T, W, Cfg::EPI_SHARED>(p, &tmAux, ew, lane, row0, col, c + 1 < nblk ? col + W : -1, v, loss_acc, wd);
                else if (row < p.M) epi_direct<W>(p, row, col, vec, v, loss_acc);
            }
            ptx::tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) { if constexpr (CG == 2) ptx::mbar_arrive_cluster(tmem_empty0 + acc * 8); else ptx::mbar_arrive(&tmem_empty[acc]); }
            if (p.epi == EPI_ATOMIC && p.out0_mc != nullptr && nblk > 0)
                mc_push_region_if_last(p, (tile * CG + (int)cta_rank) * 8 + (warp - 4), lane, row0, n0, ncols);
        }
        if (p.epi == EPI_BIAS_ACT_SE && p.loss != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
            if (lane == 0) atomicAdd(p.loss, loss_acc);
        }
    } else if (warp < kSplitWarp0) {
        // ===================================================== epilogue, 3-pass: warps 4-11, chunked promotion to registers
        if constexpr (kSplitter) ptx::setmaxnreg_inc<208>();   // 512 threads: 128 regs each at launch, 8 non-epilogue warps give 80 each
        else ptx::setmaxnreg_inc<224>();                        // 384 threads: 168 at launch, 4 non-epilogue warps give 120 each
        constexpr int HC = BN / 2;            // columns per warp: half of the tile
        const int q = warp & 3;               // TMEM lane quarter (hardware rule: warp w reads lanes 32*(w%4)..+31)
        const int half = (warp - 4) >> 2;     // column half
        const bool vec = p.vec_ok != 0;
        const bool tma = p.tma_epi != 0;
        EpiWarp ew;
        ew.aux_buf = ptx::smem_u32(epi_smem + (warp - 4) * Cfg::EPI_WARP_BYTES);
        ew.out_buf = ew.aux_buf + (Cfg::EPI_NBUF - 1) * Cfg::EPI_BLOCK_BYTES;
        ew.aux_bar = &epi_bar[warp - 4]; ew.consumed = 0; ew.in_flight = false; ew.out1_s = 1.0f;
        const uint32_t tmem_empty0 = CG == 2 ? ptx::mapa(ptx::smem_u32(&tmem_empty[0]), 0) : 0u;   // the leader's accumulator-free barriers
        int it = 0;
        float loss_acc = 0.f;
        for (int w = wid0; w < total_work; w += wstride) {
            const int tile = w % num_tiles, split = w / num_tiles;
            const int m0 = (tile / p.num_n_tiles) * (BM * CG) + (int)cta_rank * BM, n0 = (tile % p.num_n_tiles) * BN;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
            constexpr int W = Cfg::EPI_W;
            const bool pair_epi = kPresplit && EV == 0 && p.epi == EPI_BIAS_ACT_DZ && p.act == ACT_LOGISTIC && p.out1_pair && p.colsum != nullptr &&
                                  p.colsum_src == 2 && p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 && m0 + q * 32 + 32 <= p.M &&
                                  n0 + half * HC + HC <= p.N && !(p.debug & 1);
            if (tma && pair_epi) {                                                      // paired epilogue: both dA blocks of the first pair
                epi_issue_aux_pair(&tmAux, ew, lane, m0 + q * 32, n0 + half * HC);
            } else if (tma && epi_has_aux(p) && m0 + q * 32 < p.M && n0 + half * HC < p.N) {   // first aux block: requested before the K loop
                epi_issue_aux<float, W>(&tmAux, ew, lane, m0 + q * 32, n0 + half * HC);
                if (lane > 0 && lane < HC / W && n0 + half * HC + lane * W < p.N)         // the other blocks: into L2 during the K loop
                    ptx::tma_prefetch_l2_2d(&tmAux, n0 + half * HC + lane * W, m0 + q * 32);
            }
            float sum[HC];
#pragma unroll
            for (int e = 0; e < HC; ++e) sum[e] = 0.f;
            for (int kc = kb0; kc < kb1; kc += chunk_len(kc, kb0), ++it) {
                const int acc = it & 1; const uint32_t acc_ph = (it >> 1) & 1;
                if constexpr (CG == 2) ptx::mbar_wait_cluster(&tmem_full[acc], acc_ph, wd, 0x400 + acc);
                else ptx::mbar_wait(&tmem_full[acc], acc_ph, wd, 0x400 + acc);
                ptx::tcgen05_fence_after();
                const uint32_t t_row = tmem_base + acc * BN + half * HC + (static_cast<uint32_t>(q * 32) << 16);
                // drain, software-pipelined: the tcgen05.ld of group c+1 is in flight while group c is added
                uint32_t raw[2][16];
                ptx::tmem_ld_32x32b_x16(t_row, raw[0]);
#pragma unroll
                for (int c = 0; c < HC / 16; ++c) {
                    ptx::tmem_ld_wait();
                    if (c + 1 < HC / 16) ptx::tmem_ld_32x32b_x16(t_row + (c + 1) * 16, raw[(c + 1) & 1]);
#pragma unroll
                    for (int e = 0; e < 16; e += 2) ptx::add2(sum[c * 16 + e], sum[c * 16 + e + 1], raw[c & 1][e], raw[c & 1][e + 1]);   // fp32 adds, round to nearest, two per instruction
                }
                ptx::tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) { if constexpr (CG == 2) ptx::mbar_arrive_cluster(tmem_empty0 + acc * 8); else ptx::mbar_arrive(&tmem_empty[acc]); }
            }
            const int row0 = m0 + q * 32;
            const int row = row0 + lane;
            bool lean = false;
            if constexpr (kPresplit && EV == 0) {   // hot case of the forward GEMM on an interior tile: the specialised sub-block code
                lean = tma && p.epi == EPI_BIAS_ACT_DZ && p.act == ACT_LOGISTIC && p.out1_pair && p.colsum != nullptr && p.colsum_src == 2 &&
                       p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0 && row0 + 32 <= p.M && n0 + half * HC + HC <= p.N && !(p.debug & 1);
            }
            {
                float rs = p.acc_scale_ptr != nullptr ? __ldg(p.acc_scale_ptr) : 1.0f;   // undo the power-of-two operand scaling (exact)
                if (p.acc_scale_ptr2 != nullptr) rs *= __ldg(p.acc_scale_ptr2);
                if (p.row_scale != nullptr && row < p.M) {
                    const float r = __ldg(p.row_scale + row);
                    rs *= p.row_scale_inv ? __frcp_rn(r) : r;   // powers of two: exact
                }
                bool lean_mlp = false;   // EV = 1: interior tile of a hidden MLP layer (forward or dA GEMM), logistic
                if constexpr (kPresplit && EV == 1 && MA == MAJOR_K) {
                    lean_mlp = tma && p.act == ACT_LOGISTIC && row0 + 32 <= p.M && n0 + half * HC + HC <= p.N && !(p.debug & 1) &&
                               (MB == MAJOR_K ? (p.epi == EPI_BIAS_ACT && p.colsum == nullptr && p.bias != nullptr && (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0)
                                              : (p.epi == EPI_MUL_DACT && (p.colsum == nullptr || p.colsum_src == 1)));
                    lean = lean_mlp;
                }
                if (lean && (EV == 0 || MB == MAJOR_K)) {   // bias folded into the same pass: 32 independent 16-byte broadcast loads instead of 4 per sub-block
                    const float4* b4p = reinterpret_cast<const float4*>(p.bias + n0 + half * HC);
#pragma unroll
                    for (int g = 0; g < HC / 4; ++g) {
                        const float4 b4 = __ldg(b4p + g);
                        ptx::fma2(sum[g * 4], sum[g * 4 + 1], rs, b4.x, b4.y);
                        ptx::fma2(sum[g * 4 + 2], sum[g * 4 + 3], rs, b4.z, b4.w);
                    }
                } else if (p.acc_scale_ptr != nullptr || p.row_scale != nullptr) {
#pragma unroll
                    for (int e = 0; e < HC; e += 2) ptx::mul2(sum[e], sum[e + 1], rs);
                }
            }
            if (p.out1_pair) {
                ew.out1_s = p.out1_scale_ptr != nullptr ? __ldg(p.out1_scale_ptr) : 1.0f;
                if (p.out1_row_scale != nullptr && row < p.M) ew.out1_s *= __ldg(p.out1_row_scale + row);
            }
            bool lean_store = false;
            // (instantiated for the dX layout only — A K-major, B MN-major: in the variant that carries the specialised forward epilogue
            //  the extra code path costs registers — measured: 48 more bytes of spill and 8 % of the forward GEMM)
            if constexpr (kPresplit && EV == 0 && MA == MAJOR_K && MB == MAJOR_MN) {   // plain store of an interior tile (dX, gmul): stage + coalesced store, nothing else
                lean_store = tma && p.epi == EPI_STORE && p.aux0 == nullptr && p.alpha == 1.0f && p.colsum == nullptr &&
                             row0 + 32 <= p.M && n0 + half * HC + HC <= p.N && !(p.debug & 1);
                if (lean_store) {
#pragma unroll
                    for (int c = 0; c < HC / 16; ++c) {            // the two staging blocks alternate: one __syncwarp per sub-block
                        const uint32_t buf = (c & 1) ? ew.out_buf : ew.aux_buf;
                        stage_write_row<float, 16>(buf, lane, *reinterpret_cast<float (*)[16]>(&sum[c * 16]));
                        __syncwarp();
                        stage_store_interior<float, 16>(buf, lane, p.out0, p.ld_out0, row0, n0 + half * HC + c * 16);
                    }
                    __syncwarp();
                }
            }
            if constexpr (kPresplit && EV == 1 && MA == MAJOR_K) {
                if (lean) {
                    float amax = 0.f;
                    if constexpr (MB == MAJOR_K) {
                        // A = logistic(acc + b): 16-column sub-blocks through the warp's two staging blocks in turn (one __syncwarp each)
#pragma unroll
                        for (int c = 0; c < HC / 16; ++c) {
                            const uint32_t buf = (c & 1) ? ew.out_buf : ew.aux_buf;
                            float (&v)[16] = *reinterpret_cast<float (*)[16]>(&sum[c * 16]);
#pragma unroll
                            for (int e = 0; e < 16; ++e) { v[e] = act_apply(ACT_LOGISTIC, v[e]); amax = fmaxf(amax, v[e]); }
                            stage_write_row<float, 16>(buf, lane, v);
                            __syncwarp();
                            stage_store_interior<float, 16>(buf, lane, p.out0, p.ld_out0, row0, n0 + half * HC + c * 16);
                        }
                        __syncwarp();
                    } else {
                        // dZ = acc * a (1 - a), a = the layer's saved activation (aux, by TMA into aux_buf; the block after next is
                        // requested as soon as this one has been read, its latency hides behind the math and the staged store)
#pragma unroll
                        for (int c = 0; c < HC / 16; ++c) {
                            const int col = n0 + half * HC + c * 16;
                            if (!ew.in_flight) epi_issue_aux<float, 16>(&tmAux, ew, lane, row0, col);
                            ptx::mbar_wait(ew.aux_bar, ew.consumed & 1, wd, 0x600);
                            ++ew.consumed;
                            float x[16];
                            stage_read_row<float, 16>(ew.aux_buf, lane, x);
                            __syncwarp();                                          // also orders the previous sub-block's reads of out_buf
                            ew.in_flight = false;
                            if (c + 1 < HC / 16) epi_issue_aux<float, 16>(&tmAux, ew, lane, row0, col + 16);
                            float (&v)[16] = *reinterpret_cast<float (*)[16]>(&sum[c * 16]);
#pragma unroll
                            for (int e = 0; e < 16; ++e) { v[e] *= x[e] * (1.0f - x[e]); amax = fmaxf(amax, fabsf(v[e])); }
                            stage_write_row<float, 16>(ew.out_buf, lane, v);
                            __syncwarp();
                            stage_store_interior<float, 16>(ew.out_buf, lane, p.out0, p.ld_out0, row0, col);
                            if (p.colsum != nullptr) stage_colsum<float, 16>(ew.out_buf, lane, col, p.N, p.colsum);   // db of the layer below
                        }
                        __syncwarp();
                    }
                    if (p.absmax_out != nullptr) {
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
                        if (lane == 0) atomicMax(p.absmax_out, __float_as_uint(amax));
                    }
                }
            }
            if constexpr (kPresplit && EV == 0) {
                if (lean) {
#pragma unroll 1
                    for (int c = 0; c < HC / 32; ++c) {
                        const int col = n0 + half * HC + c * 32;
                        epi_fwd_pair_lean2(p, &tmAux, ew, lane, row0, col, c + 1 < HC / 32 ? col + 32 : -1, *reinterpret_cast<float (*)[32]>(&sum[0]), wd);
#pragma unroll
                        for (int e = 0; e < HC - 32; ++e) sum[e] = sum[e + 32];
                    }
                }
            }
            if (row0 < p.M && !lean && !lean_store) {
                float out_amax = 0.f;
                // ONE copy of the block code (the fused epilogue is large; unrolled four times it thrashes the instruction
                // cache): always process sum[0..31], then rotate the register accumulators down by one block.
#pragma unroll 1
                for (int c = 0; c < HC / 32; ++c) {
#pragma unroll
                    for (int u = 0; u < 32 / W; ++u) {                            // W = 16: two sub-blocks per rotation
                        const int col = n0 + half * HC + c * 32 + u * W;
                        float (&v)[W] = *reinterpret_cast<float (*)[W]>(&sum[u * W]);   // the block is processed in place ...
                        if (col < p.N) {
                            const bool more = (u + 1 < 32 / W || c + 1 < HC / 32) && col + W < p.N;
                            if (tma) epi_block<float, W, Cfg::EPI_SHARED>(p, &tmAux, ew, lane, row0, col, more ? col + W : -1, v, loss_acc, wd);
                            else if (row < p.M) epi_direct<W>(p, row, col, vec, v, loss_acc);
                            if constexpr (kPresplit && EV == 1) {   // v now holds the out0 values: their |max| saves the consumer's absmax pass
                                if (p.absmax_out != nullptr && row < p.M) {
#pragma unroll
                                    for (int e = 0; e < W; ++e) if (col + e < p.N) out_amax = fmaxf(out_amax, fabsf(v[e]));
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int e = 0; e < HC - 32; ++e) sum[e] = sum[e + 32];       // ... and the accumulators rotate afterwards
                }
                if constexpr (kPresplit && EV == 1) {
                    if (p.absmax_out != nullptr) {                                  // one red per warp and tile (fire and forget)
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) out_amax = fmaxf(out_amax, __shfl_xor_sync(0xffffffffu, out_amax, o));
                        if (lane == 0) atomicMax(p.absmax_out, __float_as_uint(out_amax));
                    }
                }
                if (p.epi == EPI_ATOMIC && p.out0_mc != nullptr && n0 + half * HC < p.N)
                    mc_push_region_if_last(p, (tile * CG + (int)cta_rank) * 8 + (warp - 4), lane, row0, n0 + half * HC, min(HC, p.N - (n0 + half * HC)));
            }
        }
        if (p.epi == EPI_BIAS_ACT_SE && p.loss != nullptr) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, o);
            if (lane == 0) atomicAdd(p.loss, loss_acc);
        }
    } else if constexpr (kSplitter) {
        // ===================================================== hi/lo splitter (3xTF32): warps 12-15
        ptx::setmaxnreg_dec<48>();
        const int t = threadIdx.x - kSplitWarp0 * 32;   // 0..127
        const uint32_t ready0 = CG == 2 ? ptx::mapa(ptx::smem_u32(&ready_bar[0]), 0) : 0u;   // the leader's `ready` barriers
        int s = 0; uint32_t ph = 0;
        for (int w = wid0; w < total_work; w += wstride) {
            const int split = w / num_tiles;
            const int kb0 = split * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.num_k_blocks);
            for (int kb = kb0; kb < kb1; ++kb) {
                ptx::mbar_wait(&full_bar[s], ph, wd, 0x500 + s);
                const uint32_t hi = ptx::smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint32_t lo = hi + Cfg::RAW_BYTES;
                if constexpr (PASSES == 2) {
                    // One task = 8 consecutive elements along the contiguous dimension of a raw fp32 tile (32 bytes) -> one 16-byte chunk
                    // of the operand's bf16 tile and one of its bf16 lo tile.
                    //  K-major raw: row r (MN index) = 32 K-elements = 128 bytes, SWIZZLE_128B: 16-byte chunk c at (c ^ (r & 7)).
                    //          bf16: rows of 64 bytes, SWIZZLE_64B: chunk c2 at (c2 ^ ((r >> 1) & 3)).
                    //  MN-major raw: boxes of [32 k-rows x 32 MN-elements (128 bytes)], SWIZZLE_128B_ATOM_32B: 32-byte unit u of row k at
                    //          (u ^ (k & 3)).   bf16: 64-element MN groups of [32 k-rows x 128 bytes], SWIZZLE_128B: chunk c at (c ^ (k & 7)).
                    // One task = 8 consecutive elements along the contiguous dimension of a raw tile (two 16-byte loads) -> one 16-byte
                    // chunk of the operand's bf16 tile and one of its bf16 lo tile.  A and B are handled by two loops so that every
                    // size / layout is a compile-time constant.
                    auto split_operand = [&](auto kmajor_tag, auto mn_tag, uint32_t raw, uint32_t t16) {
                        constexpr bool kmajor = decltype(kmajor_tag)::value;
                        constexpr int MN = decltype(mn_tag)::value;
                        constexpr uint32_t lo_off = MN * 64;                             // bytes of one bf16 tile: MN x 32 elements x 2 B
#pragma unroll 2
                        for (int ii = t; ii < MN * 4; ii += 128) {
                            uint4 x0, x1; uint32_t dst;
                            if constexpr (kmajor) {
                                const int r = ii >> 2, c2 = ii & 3;                     // MN row, 8-element K chunk
                                x0 = ptx::lds128(raw + r * 128 + (((2 * c2) ^ (r & 7)) << 4));
                                x1 = ptx::lds128(raw + r * 128 + (((2 * c2 + 1) ^ (r & 7)) << 4));
                                dst = t16 + r * 64 + ((c2 ^ ((r >> 1) & 3)) << 4);
                            } else {
                                const int k = ii & 31, m8 = ii >> 5;                    // k row, 8-element MN chunk (mn0 = 8 * m8)
                                const uint32_t src = raw + (m8 >> 2) * 4096 + k * 128 + (((m8 & 3) ^ (k & 3)) << 5);
                                // the 32 lanes of a warp read 32 different k-rows: rows k and k+4 hold the same 32-byte unit slot, so they
                                // fetch its two halves in opposite order (4-way instead of 8-way bank conflict: the optimum for 512 bytes)
                                const uint32_t flip = ((k >> 2) & 1) << 4;
                                const uint4 y0 = ptx::lds128(src + flip), y1 = ptx::lds128(src + (flip ^ 16));
                                x0 = flip ? y1 : y0; x1 = flip ? y0 : y1;
                                dst = t16 + (m8 >> 3) * 4096 + k * 128 + (((m8 & 7) ^ (k & 7)) << 4);
                            }
                            uint4 v16, l16;
                            split8_bf16(x0, x1, v16, l16);
                            ptx::sts128(dst, v16);
                            ptx::sts128(dst + lo_off, l16);
                        }
                    };
                    split_operand(std::integral_constant<bool, MA == MAJOR_K>{}, std::integral_constant<int, BM>{}, hi, lo);
                    if (!p.b_presplit)
                        split_operand(std::integral_constant<bool, MB == MAJOR_K>{}, std::integral_constant<int, BN / CG>{}, hi + Cfg::A_BYTES, lo + Cfg::A_BYTES);
                } else {
#pragma unroll 4
                for (int i = t; i < Cfg::RAW_BYTES / 16; i += 128) {
                    // kind::tf32 TRUNCATES fp32 operands (measured: tools/gemm_probe, ref_mode=1), so the raw tile already
                    // acts as hi = trunc_tf32(x); only lo = x - hi (exact in fp32) has to be materialised.
                    const uint4 x = ptx::lds128(hi + i * 16);
                    uint4 l;
                    l.x = __float_as_uint(__uint_as_float(x.x) - __uint_as_float(x.x & 0xffffe000u));
                    l.y = __float_as_uint(__uint_as_float(x.y) - __uint_as_float(x.y & 0xffffe000u));
                    l.z = __float_as_uint(__uint_as_float(x.z) - __uint_as_float(x.z & 0xffffe000u));
                    l.w = __float_as_uint(__uint_as_float(x.w) - __uint_as_float(x.w & 0xffffe000u));
                    ptx::sts128(lo + i * 16, l);
                }
                }
                ptx::fence_proxy_async_smem();
                if constexpr (CG == 2) ptx::mbar_arrive_cluster(ready0 + s * 8);
                else ptx::mbar_arrive(&ready_bar[s]);
                if (++s == STAGES) { s = 0; ph ^= 1; }
            }
        }
    }

    ptx::tcgen05_fence_before();
    __syncthreads();
    if constexpr (CG == 2) ptx::cluster_sync_all();   // neither CTA leaves (or frees TMEM) while the pair still uses its smem / barriers
    if (warp == 2) {
        ptx::tcgen05_fence_after();
        if constexpr (CG == 2) ptx::tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

}  // namespace tops
