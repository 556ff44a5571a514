// Fused-kernel instantiations for template input dimension 1 (see rff_launch.cuh).
#define TKBO_FUSED_INSTANTIATE
#include "rff_launch.cuh"

namespace tkbo {
template int fused_launch_dt<1>(int, tkbo_handle_t, tkbo_basis_t, const RffParams&, int, cudaStream_t);
}  // namespace tkbo
