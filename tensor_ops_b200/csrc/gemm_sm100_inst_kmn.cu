// tcgen05 GEMM kernel variants with A MAJOR_K, B MAJOR_MN (see gemm_sm100_launch.cuh)
#include "gemm_sm100_launch.cuh"

namespace tops {
TOPS_DEFINE_GEMM_VARIANT(gemm_launch_kmn, MAJOR_K, MAJOR_MN)
}  // namespace tops
