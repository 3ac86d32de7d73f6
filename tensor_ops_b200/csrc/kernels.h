// Launch wrappers for the bandwidth-bound (non tensor-core) kernels of libtops_b200 — internal header.
// Everything here is fp32 unless noted; all launches go to the given stream and never synchronise.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_sm100.h"

namespace tops {
namespace k {

struct LaunchCtx {
    cudaStream_t stream;
    int num_sms;
    int64_t* launches;   // incremented once per kernel launch
};

// ---- generation
void fill(const LaunchCtx&, float* p, int64_t n, float v);
void fill_bf16(const LaunchCtx&, void* p, int64_t n, float v);
void rand_normal(const LaunchCtx&, float* p, int64_t n, float mean, float sd, uint64_t seed);
void rand_uniform(const LaunchCtx&, float* p, int64_t n, float lo, float hi, uint64_t seed);
void cast_f32_bf16(const LaunchCtx&, const float* s, void* d, int64_t n);
void cast_bf16_f32(const LaunchCtx&, const void* s, float* d, int64_t n);
// v[i] = bf16(x[i]), lo[i] = bf16(x[i] - trunc_tf32(x[i]))  (operands of the bf16 correction passes, TOPS_PREC_TF32_BF16X2)
void split_bf16(const LaunchCtx&, const float* x, void* v, void* lo, int64_t n);
void eye(const LaunchCtx&, float* p, int64_t n);

// ---- F16X3 operand preparation (split_f16.cu): x -> fp16 pair (hi, lo) of x * 2^k, see the header of that file
// max |x| as fp32 bits, atomicMax into the PRE-ZEROED word out_bits
void absmax_bits(const LaunchCtx&, const float* x, int64_t n, unsigned* out_bits);
// one scale for the whole tensor, taken from absmax_bits (device); scale2 (nullable) receives {scale, 1/scale}
void split_f16_tensor(const LaunchCtx&, const float* x, int64_t n, const unsigned* absmax_bits_dev, void* hi, void* lo, float* scale2);
// the same for a [rows, cols] matrix with fp16 planes of leading dimension ld_out >= cols (padding never read)
void split_f16_tensor_2d(const LaunchCtx&, const float* x, int64_t rows, int64_t cols, int64_t ld_out, const unsigned* absmax_bits_dev, void* hi, void* lo, float* scale2);
// out[0] = 1 / (sA * sB) from two {scale, 1/scale} records
void f16x3_pair_scale(const LaunchCtx&, const float* sa2, const float* sb2, float* out);
// one scale per row: rs[r] = 1/scale_r; max_rs_bits (PRE-ZEROED) = max_r rs[r] as fp32 bits.  y (nullable, [rows, ycols]):
// max |y| is reduced into the PRE-ZEROED ymax_bits by the same launch
// A small second tensor (the layer's W) whose max|.| and fp16 pair ride along with the two launches of the row split
struct SideTensor { const float* w; int64_t n; unsigned* mx; void* hi; void* lo; float* scale2; };
// ... + fix-up pass (row-scale spread, zero rows).  Optional riders of the same two launches: `side` (see SideTensor) and the layer's
// scalars (scales_out4 = {1/sW, c, 1/c, 1/(c sW)}, from max rs, max|y| and sW2[1] — or the side tensor's scale when it rides along)
void split_f16_rows(const LaunchCtx&, const float* x, int64_t rows, int64_t cols, void* hi, void* lo, float* rs, unsigned* max_rs_bits,
                    const float* y, int64_t ycols, unsigned* ymax_bits, const SideTensor* side = nullptr, const float* sW2 = nullptr,
                    float* scales_out4 = nullptr);
// out4 = {1/sW, c, 1/c, 1/(c sW)}: the device scalars of one F16X3 ffLayer forward + VJP (sW2 = {sW, 1/sW}, nullable)
void f16x3_layer_scales(const LaunchCtx&, const unsigned* max_rs_bits, const unsigned* absmax_dA_bits, const float* sW2, float* out4);

// ---- elementwise
void axpy(const LaunchCtx&, float alpha, const float* x, const float* y /*nullable*/, float* out, int64_t n);
void add_n(const LaunchCtx&, int n_in, const float* const* xs, float* out, int64_t n);   // left fold, n_in <= 8
void sgd(const LaunchCtx&, const float* p, const float* g, float rate, float* out, int64_t n);
void dact_mul(const LaunchCtx&, int act, const float* dA, const float* A, float* dZ, int64_t n);   // dZ = dA * act'(A)
void bias_act(const LaunchCtx&, int act, const float* Z, const float* bias, float* A, int64_t rows, int64_t cols);
// lift: postfix program, up to 8 inputs
struct LiftProgram { int len; int n_consts; int32_t code[64]; float consts[16]; };
void lift(const LaunchCtx&, const LiftProgram& prog, int n_in, const float* const* in, float* out, int64_t n);
extern int64_t g_lift_catalogue_hits;   // programs served by a specialised kernel instead of the interpreter (process-wide counter)

// out_mc[i] (+)= src[i] in EVERY replica bound to the NVLS multicast address out_mc (multimem.red); n fp32 elements
void mc_push(const LaunchCtx&, const float* src, float* out_mc, int64_t n);

// ---- reductions (deterministic two-stage)
void sum_all(const LaunchCtx&, const float* x, int64_t n, float* out_scalar, float* workspace /* >= 1024 floats */);
void dot(const LaunchCtx&, const float* x, const float* y, int64_t n, float* out_scalar, float* workspace);
void trace(const LaunchCtx&, const float* a, int64_t n, int64_t ld, float* out_scalar);
// accumulate: out += column sums (out is read, elementwise, by the thread that writes it)
void col_sums(const LaunchCtx&, const float* x, int64_t rows, int64_t cols, float* out, float* workspace /* >= 64*cols floats */, bool accumulate = false);
void col_sums_bf16(const LaunchCtx&, const void* x, int64_t rows, int64_t cols, float* out, float* workspace, bool accumulate = false);

// ---- BLAS-2 / layout
void ger(const LaunchCtx&, const float* x, const float* y, float* out, int64_t n, int64_t m);
// out[n] = alpha * sum_m A(n,m) x[m] + beta*y[n];  a_tr = 0: A stored [n,m] row-major, 1: stored [m,n]
void gemv(const LaunchCtx&, float alpha, const float* a, int a_tr, const float* x, float beta, const float* y, float* out, int64_t n, int64_t m);
void transpose2d(const LaunchCtx&, const float* in, float* out, int64_t rows, int64_t cols);   // out[c,r] = in[r,c]
void permute(const LaunchCtx&, const float* in, float* out, int rank, const int64_t* in_dims, const int* perm);   // out axis a = in axis perm[a]
void broadcast_rows(const LaunchCtx&, const float* row, float* out, int64_t n, int64_t m);
void diag_embed(const LaunchCtx&, const float* v, float* out, int64_t n, int rank);
void diag_extract(const LaunchCtx&, const float* a, float* out, int64_t n, int rank);

// ---- `gmul lM 1 lN >>> sumRows` fused (x viewed as [A, R, K], y as [K, N]): out[R,N]; VJP writes dx[A,R,K] and reds into the PRE-ZEROED dy[K,N]
bool gsr_fits(int64_t K, int64_t N);
void gsr_fwd(const LaunchCtx&, const float* x, const float* y, float* out, int64_t A, int64_t R, int K, int N);
void gsr_vjp(const LaunchCtx&, const float* x, const float* y, const float* ct, float* dx, float* dy, int64_t A, int64_t R, int K, int N);
bool gsr_vjp_needs_zeroed_dy(int64_t R, int K, int N);   // false: the deterministic kernel writes every element of dy itself

// ---- losses / softmax (rows = samples)
void softmax_rows(const LaunchCtx&, const float* Z, float* A, int64_t rows, int64_t cols);   // exp / sum exp, no max-subtraction (NeuralNet.hs:52-59)
// dZ = VJP of the reference softmax TOp given dA (A not needed: recomputed from Z like the reference's closures)
void softmax_vjp_rows(const LaunchCtx&, const float* Z, const float* dA, float* dZ, int64_t rows, int64_t cols);
// fused softmax + crossEntropy head: A = softmax(Z); loss += -sum(log A * Y); dZ = VJP chain of crossEntropy∘softmax
// db (nullable, PRE-ZEROED [cols]): column sums of dZ fused in when cols <= 32 (the MNIST-style head); returns whether they were produced
bool softmax_ce_rows(const LaunchCtx&, const float* Z, const float* Y, float* A, float* dZ, float* loss, int64_t rows, int64_t cols, float* db = nullptr);
// loss VJPs on activations: squaredError dA = -2 (Y - A), loss += sum (Y-A)^2 ; crossEntropy dA = -Y / A, loss += -sum(log A * Y)
void loss_vjp(const LaunchCtx&, int loss, const float* A, const float* Y, float* dA, float* loss_out, int64_t n);

// ---- CUDA-core GEMM with the same operand/epilogue contract as the tcgen05 engine (fp32 only)
int gemm_simt(const LaunchCtx&, const GemmCall& c);
// fp32 products with one dimension <= 16 (streaming CUDA-core kernels): 1 = launched, 0 = not applicable, < 0 = -cudaError
int gemm_skinny_kind(const GemmCall& c);   // 0 = not applicable
int gemm_skinny(const LaunchCtx&, const GemmCall& c, int* colsum_fused, int* absmax_done);

}  // namespace k
}  // namespace tops
