// Standalone bring-up / tuning probe for the tcgen05 GEMM engine (not part of the product library).
//   gemm_probe dtype passes major_a major_b M N K block_n epi split_k iters [max_ctas chunk_kb no_tma cta_group]
//   dtype: 0 fp32 (passes 1/2/3), 1 bf16, 2 fp16 pairs (F16X3: the probe splits the fp32 operands itself)
// Checks the result against a double-precision reference computed on the GPU (on a row subset for big M)
// and, for fp32 single-pass, reports whether the tensor core truncates or rounds fp32 -> tf32.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../tensor_ops_b200/csrc/gemm_sm100.h"
#include "../tensor_ops_b200/csrc/kernels.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ float xform(float v, int mode) {
    if (mode == 1) return __uint_as_float(__float_as_uint(v) & 0xffffe000u);
    if (mode == 2) { unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v)); return __uint_as_float(r); }
    return v;
}

template <typename T> __device__ __forceinline__ float ldf(const T* p) { return (float)*p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

// C[r,n] = sum_k A(r,k) B(n,k) for rows r = i*row_stride
template <typename T>
__global__ void ref_gemm(const T* A, const T* B, double* C, int M, int N, int K, long long lda, long long ldb, int ma, int mb, int row_stride, int mode) {
    int n = blockIdx.x * blockDim.x + threadIdx.x;
    int ri = blockIdx.y;
    int r = ri * row_stride;
    if (n >= N || r >= M) return;
    double acc = 0;
    for (int k = 0; k < K; ++k) {
        float a = ma == 0 ? ldf(A + (long long)r * lda + k) : ldf(A + (long long)k * lda + r);
        float b = mb == 0 ? ldf(B + (long long)n * ldb + k) : ldf(B + (long long)k * ldb + n);
        acc += (double)xform(a, mode) * (double)xform(b, mode);
    }
    C[(long long)ri * N + n] = acc;
}

__global__ void fill_rand(float* p, long long n, unsigned seed, float scale) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long z = (i + 1) * 0x9E3779B97F4A7C15ull + seed * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
    p[i] = scale * (((z >> 11) * (1.0 / 9007199254740992.0)) * 2.0 - 1.0);
}
__global__ void mul_inv_scales(const float* sa2, const float* sb2, float* out) { out[0] = sa2[1] * sb2[1]; }
__global__ void to_bf16(const float* s, __nv_bfloat16* d, long long n) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) d[i] = __float2bfloat16_rn(s[i]);
}

int main(int argc, char** argv) {
    if (argc < 12) { printf("usage: gemm_probe dtype passes ma mb M N K bn epi split iters [max_ctas]\n"); return 1; }
    int dtype = atoi(argv[1]), passes = atoi(argv[2]), ma = atoi(argv[3]), mb = atoi(argv[4]);
    int M = atoi(argv[5]), N = atoi(argv[6]), K = atoi(argv[7]), bn = atoi(argv[8]), epi = atoi(argv[9]);
    int split = atoi(argv[10]), iters = atoi(argv[11]);
    int max_ctas = argc > 12 ? atoi(argv[12]) : 0;
    int chunk_kb = argc > 13 ? atoi(argv[13]) : 0;
    int no_tma = argc > 14 ? atoi(argv[14]) : 0;
    int cg = argc > 15 ? atoi(argv[15]) : 0;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    unsigned int* wd_host; unsigned int* wd_dev;
    CK(cudaHostAlloc(&wd_host, 64, cudaHostAllocMapped)); wd_host[0] = wd_host[1] = 0;
    CK(cudaHostGetDevicePointer(&wd_dev, wd_host, 0));

    const long long a_rows = ma == 0 ? M : K, a_cols = ma == 0 ? K : M;
    const long long b_rows = mb == 0 ? N : K, b_cols = mb == 0 ? K : N;
    const long long lda = a_cols, ldb = b_cols;
    float *Af, *Bf, *C, *bias;
    CK(cudaMalloc(&Af, a_rows * a_cols * 4)); CK(cudaMalloc(&Bf, b_rows * b_cols * 4));
    CK(cudaMalloc(&C, (long long)M * N * 4)); CK(cudaMalloc(&bias, N * 4));
    fill_rand<<<(a_rows * a_cols + 255) / 256, 256>>>(Af, a_rows * a_cols, 1, 1.0f);
    fill_rand<<<(b_rows * b_cols + 255) / 256, 256>>>(Bf, b_rows * b_cols, 2, 1.0f);
    fill_rand<<<(N + 255) / 256, 256>>>(bias, N, 3, 1.0f);
    void *A = Af, *B = Bf;
    __nv_bfloat16 *Ab = nullptr, *Bb = nullptr;
    if (dtype == 1) {
        CK(cudaMalloc(&Ab, a_rows * a_cols * 2)); CK(cudaMalloc(&Bb, b_rows * b_cols * 2));
        to_bf16<<<(a_rows * a_cols + 255) / 256, 256>>>(Af, Ab, a_rows * a_cols);
        to_bf16<<<(b_rows * b_cols + 255) / 256, 256>>>(Bf, Bb, b_rows * b_cols);
        A = Ab; B = Bb;
    }
    void *A2 = nullptr, *B2 = nullptr; float* acc_scale = nullptr;
    if (dtype == 2) {   // F16X3: per-tensor fp16 pairs of both operands
        unsigned* mx; float* sc; int64_t launches = 0;
        CK(cudaMalloc(&mx, 8)); CK(cudaMemset(mx, 0, 8)); CK(cudaMalloc(&sc, 5 * 4));
        void *A1, *B1;
        CK(cudaMalloc(&A1, a_rows * a_cols * 2)); CK(cudaMalloc(&A2, a_rows * a_cols * 2));
        CK(cudaMalloc(&B1, b_rows * b_cols * 2)); CK(cudaMalloc(&B2, b_rows * b_cols * 2));
        tops::k::LaunchCtx lc{0, prop.multiProcessorCount, &launches};
        cudaEvent_t s0, s1; CK(cudaEventCreate(&s0)); CK(cudaEventCreate(&s1));
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaMemset(mx, 0, 8));
            CK(cudaEventRecord(s0));
            tops::k::absmax_bits(lc, Af, a_rows * a_cols, mx);
            tops::k::split_f16_tensor(lc, Af, a_rows * a_cols, mx, A1, A2, sc);
            CK(cudaEventRecord(s1));
            tops::k::absmax_bits(lc, Bf, b_rows * b_cols, mx + 1);
            tops::k::split_f16_tensor(lc, Bf, b_rows * b_cols, mx + 1, B1, B2, sc + 2);
        }
        mul_inv_scales<<<1, 1>>>(sc, sc + 2, sc + 4);
        CK(cudaDeviceSynchronize());
        float sms; CK(cudaEventElapsedTime(&sms, s0, s1));
        float hs[5]; CK(cudaMemcpy(hs, sc, 20, cudaMemcpyDeviceToHost));
        printf("  f16 pair split of A (%lld elements): absmax + split %.4f ms = %.0f GB/s; scales A %g B %g\n", a_rows * a_cols, sms,
               8.0 * a_rows * a_cols / (sms * 1e-3) / 1e9 + 4.0 * a_rows * a_cols / (sms * 1e-3) / 1e9, hs[0], hs[2]);
        A = A1; B = B1; acc_scale = sc + 4;
    }
    CK(cudaDeviceSynchronize());

    tops::GemmCall c{};
    c.dtype = dtype; c.passes = passes; c.M = M; c.N = N; c.K = K;
    c.A = A; c.lda = lda; c.major_a = ma; c.B = B; c.ldb = ldb; c.major_b = mb;
    c.A2 = A2; c.B2 = B2; c.acc_scale_ptr = acc_scale;
    c.epi = epi; c.act = 0; c.alpha = 1.f; c.beta = 0.f; c.out0 = C; c.ld_out0 = N;
    c.bias = (epi == 2) ? bias : nullptr;
    c.split_k = split; c.block_n = bn; c.max_ctas = max_ctas; c.chunk_kb = chunk_kb; c.no_tma_epilogue = no_tma; c.cta_group = cg;
    char err[256] = {0};
    cudaStream_t st; CK(cudaStreamCreate(&st));
    auto run = [&]() {
        if (epi == 1) CK(cudaMemsetAsync(C, 0, (long long)M * N * 4, st));
        int r = tops::gemm_umma_launch(c, st, wd_dev, prop.multiProcessorCount, err, sizeof err);
        if (r != 0) { printf("LAUNCH FAIL %d: %s\n", r, err); exit(3); }
    };
    run();
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        printf("KERNEL FAIL: %s watchdog code=0x%x cta=%u\n", cudaGetErrorString(e), wd_host[0], wd_host[1]);
        return 4;
    }
    // reference on a row subset
    int row_stride = M > 4096 ? (M / 2048) | 1 : 1;   // odd stride: hits every lane position inside tiles
    int nref = (M + row_stride - 1) / row_stride;
    double* Cref; CK(cudaMalloc(&Cref, (long long)nref * N * 8));
    std::vector<float> hC((long long)M * N);
    CK(cudaMemcpy(hC.data(), C, (long long)M * N * 4, cudaMemcpyDeviceToHost));
    std::vector<float> hb(N);
    CK(cudaMemcpy(hb.data(), bias, N * 4, cudaMemcpyDeviceToHost));
    std::vector<double> hR((long long)nref * N);
    const int nmodes = (dtype == 0 && passes == 1) ? 3 : 1;
    for (int mode = 0; mode < nmodes; ++mode) {
        dim3 g((N + 127) / 128, nref);
        if (dtype == 1) ref_gemm<__nv_bfloat16><<<g, 128>>>(Ab, Bb, Cref, M, N, K, lda, ldb, ma, mb, row_stride, mode);
        else ref_gemm<float><<<g, 128>>>(Af, Bf, Cref, M, N, K, lda, ldb, ma, mb, row_stride, mode);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(hR.data(), Cref, (long long)nref * N * 8, cudaMemcpyDeviceToHost));
        double num = 0, den = 0, maxabs = 0;
        for (int ri = 0; ri < nref; ++ri) {
            long long r = (long long)ri * row_stride;
            for (int n = 0; n < N; ++n) {
                double ref = hR[(long long)ri * N + n];
                if (epi == 2) ref += hb[n];
                double d = (double)hC[r * N + n] - ref;
                num += d * d; den += ref * ref; if (fabs(d) > maxabs) maxabs = fabs(d);
            }
        }
        printf("  ref_mode=%d (0 exact,1 trunc,2 rna): rel_fro_err=%.3e max_abs_err=%.3e\n", mode, sqrt(num / (den + 1e-300)), maxabs);
    }
    // timing
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) run();
    CK(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) run();
    CK(cudaEventRecord(e1, st));
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { printf("KERNEL FAIL (timing): %s watchdog code=0x%x\n", cudaGetErrorString(e), wd_host[0]); return 4; }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    printf("RESULT dtype=%d passes=%d ma=%d mb=%d M=%d N=%d K=%d bn=%d epi=%d split=%d ctas=%d chunk=%d notma=%d cg=%d : %.4f ms  %.1f TFLOP/s\n",
           dtype, passes, ma, mb, M, N, K, bn, epi, split, max_ctas, chunk_kb, no_tma, cg, ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
    return 0;
}
