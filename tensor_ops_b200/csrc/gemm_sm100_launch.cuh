// Kernel instantiation + launch templates of the tcgen05 GEMM, shared by the per-layout translation units
// (gemm_sm100_inst_*.cu): the 64 kernel variants are split over four files by operand layout so that they compile in parallel.
#pragma once

#include <cuda_runtime.h>

#include "gemm_sm100.cuh"

namespace tops {

template <typename T, int MA, int MB, int BN, int STAGES, int PASSES, int CG, int EV = 0>
cudaError_t launch_one(const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
    using Cfg = GemmCfg<T, MA, MB, BN, STAGES, PASSES, CG>;
    auto kern = gemm_umma_kernel<T, MA, MB, BN, STAGES, PASSES, CG, EV>;
    // opt in to > 48 KiB of dynamic shared memory: once per (instantiation, device) — the attribute is per device
    static unsigned long long attr_done_mask = 0;   // bit d = done on device d (racing threads at worst repeat the idempotent call)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 64 || !((attr_done_mask >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        if (dev < 64) attr_done_mask |= 1ull << dev;
    }
    if constexpr (CG == 1) {
        kern<<<grid, Cfg::NUM_THREADS, Cfg::SMEM_BYTES, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], p);
        return cudaGetLastError();
    } else {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(Cfg::NUM_THREADS); cfg.dynamicSmemBytes = Cfg::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, kern, tm[0], tm[1], tm[2], tm[3], tm[4], p);
    }
}

template <typename T, int MA, int MB, int CG>
cudaError_t launch_cg(int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
    // stage bytes: 1-pass (128 + bn/CG) * 128, 3-pass twice that; stage counts fill the 227 KiB left after the epilogue staging
    if constexpr (sizeof(T) == 4) {
        if (passes == 2) {   // TF32 hi*hi + two bf16 correction passes
            if (bn == 128) return launch_one<T, MA, MB, 128, CG == 2 ? 4 : 3, 2, CG>(tm, p, grid, st);
            return launch_one<T, MA, MB, 256, CG == 2 ? 3 : 2, 2, CG>(tm, p, grid, st);
        }
        if (passes >= 2) {
            if (bn == 128) return launch_one<T, MA, MB, 128, CG == 2 ? 4 : 3, 3, CG>(tm, p, grid, st);
            return launch_one<T, MA, MB, 256, CG == 2 ? 3 : 2, 3, CG>(tm, p, grid, st);
        }
        // fp32 1-pass: one stage fewer than fits, so that each epilogue warp gets separate aux and output staging blocks
        if (bn == 128) return launch_one<T, MA, MB, 128, CG == 2 ? 6 : 5, 1, CG>(tm, p, grid, st);
        return launch_one<T, MA, MB, 256, CG == 2 ? 5 : 3, 1, CG>(tm, p, grid, st);
    } else if constexpr (std::is_same<T, __half>::value) {   // F16X3: (hi, lo) planes of both operands per stage: 2 * (128 + bn/CG) * 128 bytes
        if constexpr (MA == MAJOR_K && CG == 2) {   // passes == 5: the MLP epilogue variant (EV = 1) of the 256-column pair kernel
            if (passes == 5 && bn == 256) return launch_one<T, MA, MB, 256, 3, 4, CG, 1>(tm, p, grid, st);
        }
        if (bn == 128) return launch_one<T, MA, MB, 128, CG == 2 ? 4 : 3, 4, CG>(tm, p, grid, st);
        return launch_one<T, MA, MB, 256, CG == 2 ? 3 : 2, 4, CG>(tm, p, grid, st);
    } else {
        if (bn == 128) return launch_one<T, MA, MB, 128, CG == 2 ? 8 : 6, 1, CG>(tm, p, grid, st);
        return launch_one<T, MA, MB, 256, CG == 2 ? 6 : 4, 1, CG>(tm, p, grid, st);
    }
}

template <typename T, int MA, int MB>
cudaError_t launch_major(int cg, int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {
    if (cg == 2) return launch_cg<T, MA, MB, 2>(bn, passes, tm, p, grid, st);
    return launch_cg<T, MA, MB, 1>(bn, passes, tm, p, grid, st);
}

// one entry per (A layout, B layout); dtype 0 = fp32 operands, 1 = bf16, 2 = fp16 pairs (F16X3)
#define TOPS_DEFINE_GEMM_VARIANT(NAME, MA, MB)                                                                                          \
    cudaError_t NAME(int dtype, int cg, int bn, int passes, const CUtensorMap* tm, const GemmParams& p, int grid, cudaStream_t st) {    \
        if (dtype == 1) return launch_major<__nv_bfloat16, MA, MB>(cg, bn, 1, tm, p, grid, st);                                        \
        if (dtype == 2) return launch_major<__half, MA, MB>(cg, bn, passes, tm, p, grid, st);                                               \
        return launch_major<float, MA, MB>(cg, bn, passes, tm, p, grid, st);                                                           \
    }

}  // namespace tops
