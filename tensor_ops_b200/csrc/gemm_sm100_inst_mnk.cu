// tcgen05 GEMM kernel variants with A MAJOR_MN, B MAJOR_K (see gemm_sm100_launch.cuh)
#include "gemm_sm100_launch.cuh"

namespace tops {
TOPS_DEFINE_GEMM_VARIANT(gemm_launch_mnk, MAJOR_MN, MAJOR_K)
}  // namespace tops
