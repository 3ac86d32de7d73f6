// TMA tensor-map construction, tiling / split-K decisions and dispatch for the tcgen05 GEMM (gemm_sm100.cuh); the kernel
// instantiations live in gemm_sm100_inst_*.cu.
#include "gemm_sm100.h"

#include <cuda.h>
#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "gemm_sm100.cuh"

namespace tops {

namespace {

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_encode_once;

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    std::call_once(g_encode_once, [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    });
    return g_encode;
}

// 2-D row-major tensor [outer, inner] with leading dimension ld (elements); box = box_inner x box_outer, 128B swizzle.
enum Swz { SWZ_128 = 0, SWZ_128_ATOM32 = 1, SWZ_64 = 2 };
bool make_map(CUtensorMap* m, int dtype, const void* ptr, long long inner, long long outer, long long ld, int box_inner, int box_outer, int swz) {
    auto enc = get_encode();
    if (!enc) return false;
    const size_t es = dtype == 0 ? 4 : 2;   // 0 fp32, 1 bf16, 2 fp16
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) return false;
    if (((size_t)ld * es) % 16 != 0) return false;
    if (inner <= 0 || outer <= 0) return false;
    cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, dtype == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : dtype == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                     const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     swz == SWZ_128_ATOM32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : swz == SWZ_64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace

// defined in gemm_sm100_inst_{kk,kmn,mnk,mnmn}.cu
cudaError_t gemm_launch_kk(int dtype, int cg, int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st);
cudaError_t gemm_launch_kmn(int dtype, int cg, int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st);
cudaError_t gemm_launch_mnk(int dtype, int cg, int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st);
cudaError_t gemm_launch_mnmn(int dtype, int cg, int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st);

namespace {

cudaError_t launch_any(int dtype, int cg, int ma, int mb, int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
    if (ma == MAJOR_K && mb == MAJOR_K) return gemm_launch_kk(dtype, cg, bn, passes, tm, p, grid, st);
    if (ma == MAJOR_K && mb == MAJOR_MN) return gemm_launch_kmn(dtype, cg, bn, passes, tm, p, grid, st);
    if (ma == MAJOR_MN && mb == MAJOR_K) return gemm_launch_mnk(dtype, cg, bn, passes, tm, p, grid, st);
    return gemm_launch_mnmn(dtype, cg, bn, passes, tm, p, grid, st);
}

}  // namespace

int gemm_umma_launch(const GemmCall& c, cudaStream_t stream, unsigned int* watchdog_dev, int num_sms, char* err, size_t errlen) {
    auto fail = [&](int code, const char* msg) {
        if (err && errlen) snprintf(err, errlen, "%s", msg);
        return code;
    };
    if (c.M <= 0 || c.N <= 0 || c.K <= 0) return fail(-1, "gemm: empty problem");
    if (c.passes >= 2 && c.dtype != 0) return fail(-1, "gemm: the split modes need fp32 operands");
    if (c.dtype == 2 && (!c.A2 || !c.B2)) return fail(-1, "gemm: fp16-pair operands need both planes");
    if (c.dtype == 2 && c.io_bf16) return fail(-1, "gemm: the fp16-pair mode writes fp32 outputs");
    int passes = c.dtype == 2 ? 4 : c.dtype == 1 ? 1 : c.passes >= 2 && c.passes <= 3 ? c.passes : 1;
    const int es = c.dtype == 0 ? 4 : 2;
    const int kb_elems = 128 / es;
    int bn = c.block_n;
    if (bn == 0) bn = (c.N <= 128) ? 128 : 256;
    if (bn != 128 && bn != 256) return fail(-1, "gemm: block_n must be 128 or 256");

    // CTA pairs (cta_group::2): a 2-CTA cluster computes a 256-row tile, each CTA staging its 128 rows of A and half of the B
    // tile — per-SM shared-memory traffic per MMA drops from 12 KiB to 8 KiB, which is what bounds the 1-CTA kernels.
    int cg = c.cta_group;
    if (cg == 0) cg = (c.M > 128) ? 2 : 1;
    if (cg != 1 && cg != 2) return fail(-1, "gemm: cta_group must be 0 (auto), 1 or 2");
    CUtensorMap tm[5];   // A, B, aux0, bf16(B), bf16(B_lo)
    CUtensorMap &ta = tm[0], &tb = tm[1];
    memset(tm, 0, sizeof tm);
    bool ok;
    if (c.major_a == MAJOR_K) ok = make_map(&ta, c.dtype, c.A, c.K, c.M, c.lda, kb_elems, 128, SWZ_128);
    else ok = make_map(&ta, c.dtype, c.A, c.M, c.K, c.lda, kb_elems, kb_elems, c.dtype == 0 ? SWZ_128_ATOM32 : SWZ_128);
    if (!ok) return fail(-1, "gemm: operand A not expressible as a TMA tensor map (alignment/stride)");
    if (c.major_b == MAJOR_K) ok = make_map(&tb, c.dtype, c.B, c.K, c.N, c.ldb, kb_elems, bn / cg, SWZ_128);
    else ok = make_map(&tb, c.dtype, c.B, c.N, c.K, c.ldb, kb_elems, kb_elems, c.dtype == 0 ? SWZ_128_ATOM32 : SWZ_128);
    if (!ok) return fail(-1, "gemm: operand B not expressible as a TMA tensor map (alignment/stride)");
    if (c.dtype == 2) {   // lo planes: same geometry as the hi planes, delivered into the second half of every stage
        if (c.major_a == MAJOR_K) ok = make_map(&tm[3], 2, c.A2, c.K, c.M, c.lda, kb_elems, 128, SWZ_128);
        else ok = make_map(&tm[3], 2, c.A2, c.M, c.K, c.lda, kb_elems, kb_elems, SWZ_128);
        if (ok) {
            if (c.major_b == MAJOR_K) ok = make_map(&tm[4], 2, c.B2, c.K, c.N, c.ldb, kb_elems, bn / cg, SWZ_128);
            else ok = make_map(&tm[4], 2, c.B2, c.N, c.K, c.ldb, kb_elems, kb_elems, SWZ_128);
        }
        if (!ok) return fail(-1, "gemm: lo planes not expressible as TMA tensor maps (alignment/stride)");
    }
    // Staged epilogue: outputs are transposed through shared memory and written with coalesced 16-byte stores, the aux operand
    // (if any) is fetched by TMA as 32 x 32 boxes (fp32: 128-byte rows, SWIZZLE_128B; bf16: 64-byte rows, SWIZZLE_64B).
    // Needs 16-byte aligned rows everywhere; otherwise the kernel falls back to its direct register<->global epilogue.
    bool tma_epi = c.epi != EPI_ATOMIC && !c.no_tma_epilogue && (c.io_bf16 != 0) == (c.dtype == 1);   // staged blocks have the operand dtype
    if (tma_epi) {
        const int oes = c.io_bf16 ? 2 : 4;
        auto aligned = [&](const void* ptr, long long ld) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((size_t)ld * oes) % 16 == 0; };
        const bool has_out1 = c.epi == EPI_BIAS_ACT_DZ || c.epi == EPI_BIAS_ACT_SE;
        const bool has_aux = has_out1 || c.epi == EPI_MUL_DACT || (c.epi == EPI_STORE && c.aux0 != nullptr);
        tma_epi = aligned(c.out0, c.ld_out0);
        if (tma_epi && has_out1 && c.out1_pair)   // fp16 planes: 16-byte aligned rows of 2-byte elements
            tma_epi = c.out1b && (reinterpret_cast<uintptr_t>(c.out1) & 15) == 0 && (reinterpret_cast<uintptr_t>(c.out1b) & 15) == 0 && ((size_t)c.ld_out1 * 2) % 16 == 0;
        else if (tma_epi && has_out1) tma_epi = aligned(c.out1, c.ld_out1);
        // aux boxes: 32 rows x 32 columns (fp32: 128-byte rows, bf16: 64-byte rows); F16X3 works on 32 x 16 fp32 sub-blocks (64-byte rows)
        if (tma_epi && has_aux) tma_epi = make_map(&tm[2], c.io_bf16 ? 1 : 0, c.aux0, c.N, c.M, c.ld_aux0, c.dtype == 2 ? 16 : 32, 32, (c.io_bf16 || c.dtype == 2) ? SWZ_64 : SWZ_128);
    }
    if (c.out1_pair && !tma_epi) return fail(-1, "gemm: the fp16-pair output needs the staged epilogue (16-byte aligned rows)");
    if ((c.acc_scale_ptr || c.row_scale || c.out1_pair) && passes < 2) return fail(-1, "gemm: accumulator scaling / pair output exist in the chunked kernels only");
    // passes == 2 with a pre-split B: two more maps delivering the bf16 tiles in the layouts the splitter would have produced
    // (K-major: 64-byte rows, SWIZZLE_64B; MN-major: [32 k-rows x 64 elements], SWIZZLE_128B)
    bool b_presplit = false;
    if (passes == 2 && c.B16 && c.Blo16) {
        if (c.major_b == MAJOR_K) b_presplit = make_map(&tm[3], 1, c.B16, c.K, c.N, c.ldb, 32, bn / cg, SWZ_64) && make_map(&tm[4], 1, c.Blo16, c.K, c.N, c.ldb, 32, bn / cg, SWZ_64);
        else b_presplit = make_map(&tm[3], 1, c.B16, c.N, c.K, c.ldb, 64, 32, SWZ_128) && make_map(&tm[4], 1, c.Blo16, c.N, c.K, c.ldb, 64, 32, SWZ_128);
    }
    GemmParams p{};
    p.b_presplit = b_presplit ? 1 : 0;
    p.M = c.M; p.N = c.N; p.K = c.K;
    p.num_m_tiles = (c.M + 128 * cg - 1) / (128 * cg);
    p.num_n_tiles = (c.N + bn - 1) / bn;
    p.num_k_blocks = (c.K + kb_elems - 1) / kb_elems;
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int ctas = (c.max_ctas > 0 ? c.max_ctas : num_sms) / cg;   // CTA groups that can be resident
    int split = c.split_k;
    if (c.epi != EPI_ATOMIC) split = 1;
    else if (split <= 0) {
        split = 1;
        if (tiles < ctas) {
            split = (2 * ctas) / tiles;
            const int max_split = p.num_k_blocks / 8 > 0 ? p.num_k_blocks / 8 : 1;   // keep >= 8 k-blocks per work item
            if (split > max_split) split = max_split;
            if (split < 1) split = 1;
        }
    }
    if (split > p.num_k_blocks) split = p.num_k_blocks;
    p.kb_per_split = (p.num_k_blocks + split - 1) / split;
    p.split_k = (p.num_k_blocks + p.kb_per_split - 1) / p.kb_per_split;
    // 3-pass: TMEM accumulation truncates (~2e-8 relative per MMA, linear in the chain length), so a chunk of
    // chunk_kb k-blocks (12 MMAs each) is promoted to round-to-nearest fp32 register sums; 4 k-blocks = 128 K-elements.
    // F16X3: k-blocks hold 64 K-elements and 12 MMAs; 2 k-blocks = 128 K-elements = 24 MMAs per chunk.
    p.chunk_kb = c.chunk_kb > 0 ? c.chunk_kb : (c.dtype == 2 ? 2 : 4);
    p.chunk_head_kb = c.chunk_head_kb > p.chunk_kb ? c.chunk_head_kb : p.chunk_kb;
    p.acc_scale_ptr = c.acc_scale_ptr; p.acc_scale_ptr2 = c.acc_scale_ptr ? c.acc_scale_ptr2 : nullptr; p.row_scale = c.row_scale; p.row_scale_inv = c.row_scale_inv;
    p.out1_pair = c.out1_pair; p.out1b = c.out1b; p.out1_scale_ptr = c.out1_scale_ptr; p.out1_row_scale = c.out1_row_scale;
    p.epi = c.epi; p.act = c.act; p.alpha = c.alpha; p.beta = c.beta;
    p.out0 = c.out0; p.ld_out0 = c.ld_out0;
    p.out1 = c.out1; p.ld_out1 = c.ld_out1;
    p.aux0 = c.aux0; p.ld_aux0 = c.ld_aux0;
    p.bias = c.bias; p.loss = c.loss;
    p.io_bf16 = c.io_bf16;
    p.watchdog = watchdog_dev;
    { static const int dbg = getenv("TOPS_GEMM_DEBUG") ? atoi(getenv("TOPS_GEMM_DEBUG")) : 0; p.debug = dbg; }
    p.tma_epi = tma_epi ? 1 : 0;
    // fused column sums ride on the staged blocks of the TMA epilogue; tell the caller whether they were produced
    p.colsum = (tma_epi && c.colsum_src != 0) ? c.colsum : nullptr;
    p.colsum_src = c.colsum_src;
    p.out0_mc = c.epi == EPI_ATOMIC ? c.out0_mc : nullptr; p.tile_counters = c.tile_counters;
    if (p.out0_mc && !p.tile_counters) return fail(-1, "gemm: out0_mc needs tile_counters");
    if (c.colsum_fused) *c.colsum_fused = p.colsum != nullptr ? 1 : 0;
    auto vec_ok = [&](const void* ptr, long long ld) {
        if (!ptr) return true;
        const int oes = c.io_bf16 ? 2 : 4;
        return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ((size_t)ld * oes) % 16 == 0;
    };
    p.vec_ok = vec_ok(c.out0, c.ld_out0) && vec_ok(c.out1, c.ld_out1) && vec_ok(c.aux0, c.ld_aux0);
    if (c.epi == EPI_ATOMIC) p.vec_ok = (reinterpret_cast<uintptr_t>(c.out0) & 15) == 0 && (c.ld_out0 % 4) == 0;

    const int total = tiles * p.split_k;
    const int grid = (total < ctas ? total : ctas) * cg;
    cudaError_t e;
    // the MLP's hidden-layer GEMMs (forward / dA) have their own epilogue variant of the fp16-pair kernel (EV = 1): specialised
    // interior-tile code for logistic, and max|out0| as a side output for every activation
    if (c.dtype == 2 && cg == 2 && bn == 256 && c.major_a == MAJOR_K &&
        ((c.epi == EPI_BIAS_ACT && c.major_b == MAJOR_K) || (c.epi == EPI_MUL_DACT && c.major_b == MAJOR_MN)))
        passes = 5;
    p.absmax_out = passes == 5 ? c.absmax_out : nullptr;
    if (c.absmax_done) *c.absmax_done = p.absmax_out != nullptr ? 1 : 0;
    e = launch_any(c.dtype, cg, c.major_a, c.major_b, bn, passes, tm, p, grid, stream);
    if (e != cudaSuccess) {
        if (err && errlen) snprintf(err, errlen, "gemm launch: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

}  // namespace tops
