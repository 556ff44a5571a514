// Launch of one instantiation of the fused kernel.  The instantiations are compiled in one translation unit per template
// input dimension (rff_inst_dt*.cu, built in parallel); rff.cu only sees fused_launch_dt<DT>.
#pragma once
#include "common.h"
#include "rff_fused.cuh"

namespace tkbo {

// all MODE / NPG / precision / pair variants of one input dimension; MODE 2 (duel) needs an even d
template <int DT>
int fused_launch_dt(int mode, tkbo_handle_t h, tkbo_basis_t bs, const RffParams& p, int grid, cudaStream_t stream);

#ifdef TKBO_FUSED_INSTANTIATE
template <int DT, int MODE, int NPG, bool C8>
static int launch_fused_c(tkbo_handle_t h, tkbo_basis_t bs, const RffParams& p, int grid, cudaStream_t stream) {
  if constexpr (NPG != 4) {
    if (bs->pair) {
      // clusters of two CTAs (one TPC): an even grid, one item stream per pair
      auto kern = rff_fused_kernel<DT, MODE, NPG, C8, true>;
      TKBO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bs->smem_bytes));
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(static_cast<unsigned>(grid), 1, 1);
      cfg.blockDim = dim3(kNumThreadsPair(NPG), 1, 1);
      cfg.dynamicSmemBytes = static_cast<size_t>(bs->smem_bytes);
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      TKBO_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, bs->tm_hi, bs->tm_lo, bs->tm_w, p));
      h->launches += 1;
      return TKBO_OK;
    }
  }
  auto kern = rff_fused_kernel<DT, MODE, NPG, C8>;
  TKBO_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bs->smem_bytes));
  kern<<<grid, kNumThreadsFor(NPG), bs->smem_bytes, stream>>>(bs->tm_hi, bs->tm_lo, bs->tm_w, p);
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

template <int DT, int MODE, int NPG>
static int launch_fused_n(tkbo_handle_t h, tkbo_basis_t bs, const RffParams& p, int grid, cudaStream_t stream) {
  return bs->corr8 ? launch_fused_c<DT, MODE, NPG, true>(h, bs, p, grid, stream)
                   : launch_fused_c<DT, MODE, NPG, false>(h, bs, p, grid, stream);
}

template <int DT, int MODE>
static int launch_fused(tkbo_handle_t h, tkbo_basis_t bs, const RffParams& p, int grid, cudaStream_t stream) {
  return bs->npg == 4 ? launch_fused_n<DT, MODE, 4>(h, bs, p, grid, stream)
         : bs->npg == 3 ? launch_fused_n<DT, MODE, 3>(h, bs, p, grid, stream)
                        : launch_fused_n<DT, MODE, 2>(h, bs, p, grid, stream);
}


template <int DT>
int fused_launch_dt(int mode, tkbo_handle_t h, tkbo_basis_t bs, const RffParams& p, int grid, cudaStream_t stream) {
  if (mode == 0) return launch_fused<DT, 0>(h, bs, p, grid, stream);
  if (mode == 1) return launch_fused<DT, 1>(h, bs, p, grid, stream);
  if constexpr (DT % 2 == 0) {
    if (mode == 2) return launch_fused<DT, 2>(h, bs, p, grid, stream);
  }
  set_error("internal: unsupported (DT, MODE)");
  return TKBO_ERR_UNSUPPORTED;
}
#endif

}  // namespace tkbo
