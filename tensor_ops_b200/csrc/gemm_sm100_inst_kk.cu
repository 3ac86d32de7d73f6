// tcgen05 GEMM kernel variants with A MAJOR_K, B MAJOR_K (see gemm_sm100_launch.cuh)
#include "gemm_sm100_launch.cuh"

namespace tops {
TOPS_DEFINE_GEMM_VARIANT(gemm_launch_kk, MAJOR_K, MAJOR_K)
}  // namespace tops
