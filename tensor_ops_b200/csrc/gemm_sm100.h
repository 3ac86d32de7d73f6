// Host-side interface of the tcgen05 GEMM engine (internal to libtops_b200; the public C ABI is include/tops_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

namespace tops {

struct GemmCall {
    int dtype;      // 0 = fp32 operands (TF32 tensor cores), 1 = bf16 operands, 2 = fp16 PAIRS (F16X3: A/B = hi planes, A2/B2 = lo planes,
                    //     lda/ldb in fp16 elements; three fp16 passes hi*hi + lo*hi + hi*lo, fp32 accumulate; outputs / aux are fp32)
    int passes;     // dtype 0: 1 = single TF32 pass, 2 = TF32 + two bf16 correction passes, 3 = 3xTF32 hi/lo split; ignored otherwise
    int M, N, K;
    const void* A; long long lda; int major_a;   // 0: A stored [M,K] (K contiguous), 1: stored [K,M]
    const void* B; long long ldb; int major_b;   // 0: B stored [N,K] (K contiguous), 1: stored [K,N]
    const void* B16; const void* Blo16;   // optional (passes == 2): bf16(B) and bf16(B - trunc_tf32(B)), same shape / ldb as B, pre-split in HBM
    const void* A2; const void* B2;       // dtype 2: the lo planes (same shape / leading dimension as A / B)
    // dtype 2: the operands were scaled by powers of two when they were split; the accumulators are multiplied by
    // (*acc_scale_ptr) * row_scale[m] (or / row_scale[m]) before the epilogue math (all factors powers of two: exact)
    const float* acc_scale_ptr;           // device scalar, NULL = 1
    const float* acc_scale_ptr2;          // second device scalar factor (needs acc_scale_ptr), NULL = 1
    const float* row_scale; int row_scale_inv;   // [M] device vector, NULL = 1
    // out1 as an fp16 pair for the next F16X3 GEMM: t = out1 * (*out1_scale_ptr) * out1_row_scale[m]; out1 <- fp16(t), out1b <- fp16(t - fp16(t))
    int out1_pair; void* out1b; const float* out1_scale_ptr; const float* out1_row_scale;
    unsigned int* absmax_out;   // dtype 2, BIAS_ACT / MUL_DACT: atomicMax of the float bits of |out0| (zeroed by the caller); NULL = none
    int* absmax_done;           // out: 1 when the launched variant produced absmax_out (pair kernel, 256-column tiles), else 0
    int epi, act;
    float alpha, beta;
    void* out0; long long ld_out0;
    void* out1; long long ld_out1;
    const void* aux0; long long ld_aux0;
    const float* bias;
    float* loss;
    int io_bf16;
    int split_k;    // 0 = choose automatically (only EPI_ATOMIC may split)
    int block_n;    // 0 = choose automatically, else 128 or 256
    int max_ctas;   // 0 = number of SMs
    int cta_group;  // 0 = choose automatically, 1 = one CTA per tile, 2 = CTA pairs (tcgen05 cta_group::2, 256-row tiles)
    int accumulate; // EPI_ATOMIC only: add into out0 as it is (the caller zeroed it / is accumulating across calls)
    float* colsum;         // optional [N] fp32 accumulator, PRE-ZEROED by the caller: column sums of out0 (colsum_src 1) / out1 (2)
    int colsum_src;        // 0 = none
    int* colsum_fused;     // out: set to 1 if the kernel produced the column sums (TMA epilogue), else 0 (caller reduces separately)
    float* out0_mc;        // EPI_ATOMIC only: NVLS multicast alias of a buffer shaped like out0; finished split-K regions are pushed there once
    int* tile_counters;    // with out0_mc: zeroed int[ceil(M/128) * ceil(N/bn) * 8 + 16] region-arrival counters
    int no_tma_epilogue;   // 1 = force the direct register<->global epilogue (bring-up / A-B comparison)
    int chunk_kb;   // chunked kernels: k-blocks per TMEM chunk before promotion to fp32 registers (0 = default: 4, F16X3 2)
    int chunk_head_kb;   // chunked kernels: k-blocks in each of the first two chunks of a work item (0 = chunk_kb)
    const char* tag;   // profiling label (tops_profile_*), may be NULL
};

// returns 0 on success; -1 if the problem cannot be expressed as TMA tensor maps (caller uses the SIMT kernel);
// >0 cudaError_t on launch failure.  `err` receives a message.
int gemm_umma_launch(const GemmCall& c, cudaStream_t stream, unsigned int* watchdog_dev, int num_sms, char* err, size_t errlen);

}  // namespace tops
