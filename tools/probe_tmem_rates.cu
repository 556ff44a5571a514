// Micro-probe (B200): per-SM throughput of tcgen05.ld / tcgen05.st (32x32b.x32), broadcast LDS.128 / LDS.64 / LDS.32
// and MUFU.COS, with 4 / 8 / 16 warps.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe probe_tmem_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../top-k-ranking-bayesian-optimization_b200/csrc/ptx.cuh"
using namespace tkbo;

__global__ void probe(int mode, int iters, unsigned long long* out_cycles, float* sink) {
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(16) float sh[1024];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = i * 0.001f;
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t taddr = tmem_base_s + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 32 % 448;
  uint32_t v[32];
  for (int i = 0; i < 32; ++i) v[i] = lane + i;
  float acc = 0.f;
  __syncthreads();
  unsigned long long t0 = clock64();
  if (mode == 0) {
    for (int it = 0; it < iters; ++it) { tmem_ld_x32(taddr, v); tmem_wait_ld(); acc += __uint_as_float(v[it & 31]); }
  } else if (mode == 1) {
    for (int it = 0; it < iters; ++it) { v[0] += it; tmem_st_x32(taddr, v); }
    tmem_wait_st();
  } else if (mode == 2) {
    for (int it = 0; it < iters; ++it) { float4 w = *reinterpret_cast<const float4*>(&sh[(it * 4) & 1023]); acc += w.x + w.y + w.z + w.w; }
  } else if (mode == 3) {
    for (int it = 0; it < iters; ++it) { float2 w = *reinterpret_cast<const float2*>(&sh[(it * 2) & 1023]); acc += w.x + w.y; }
  } else if (mode == 4) {
    for (int it = 0; it < iters; ++it) { acc += sh[it & 1023]; }
  } else if (mode == 5) {
    float x = lane * 0.01f;
#pragma unroll 8
    for (int it = 0; it < iters; ++it) { acc += __cosf(x + it); }
  } else if (mode == 6) {   // ld without wait each time (pipelined)
    for (int it = 0; it < iters; it += 4) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld_x32(taddr, a); tmem_ld_x32(taddr, b); tmem_ld_x32(taddr, c); tmem_ld_x32(taddr, d); tmem_wait_ld();
      acc += __uint_as_float(a[1] ^ b[2] ^ c[3] ^ d[4]);
    }
  }
  __syncthreads();
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base_s, 512);
}

int main() {
  unsigned long long* d_c; float* d_s;
  cudaMalloc(&d_c, 8 * 1024); cudaMalloc(&d_s, 4 * 1024 * 1024);
  const char* names[] = {"LDTM.x32+wait", "STTM.x32", "LDS.128 bcast", "LDS.64 bcast", "LDS.32 bcast", "MUFU.COS", "LDTM.x32 x4 pipelined"};
  for (int mode = 0; mode < 7; ++mode)
    for (int warps : {4, 8, 16}) {
      const int iters = 4096;
      probe<<<148, warps * 32>>>(mode, iters, d_c, d_s);
      cudaDeviceSynchronize();
      probe<<<148, warps * 32>>>(mode, iters, d_c, d_s);
      cudaError_t e = cudaDeviceSynchronize();
      unsigned long long c[148];
      cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
      double cyc = (double)c[0];
      double per_warp_inst = cyc / iters;                         // cycles per iteration (all warps concurrently)
      double sm_inst_per_cyc = warps * (double)iters / cyc;        // warp-instructions per cycle per SM
      printf("%-24s warps=%2d  %.2f cyc/iter  %.3f warp-inst/cyc/SM  (%s)\n", names[mode], warps, per_warp_inst, sm_inst_per_cyc,
             cudaGetErrorString(e));
    }
  return 0;
}
