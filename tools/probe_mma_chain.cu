// Micro-probe (B200): cycles per tcgen05.mma (kind::f16, M=128, A from TMEM, B from smem) as a function of N, of the
// number of MMAs issued per elected region, and of how the issuing thread is selected.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_mma_chain probe_mma_chain.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../top-k-ranking-bayesian-optimization_b200/csrc/ptx.cuh"
using namespace tkbo;

// mode 0: elect_one + __syncwarp around a group of `per` MMAs;  mode 1: lane 0 runs the whole loop alone;
// mode 2: like 0 but the loop-carried addresses are forced uniform with __shfl_sync(…, 0)
template <int PER>
__global__ void probe(int N, int mode, int groups, unsigned long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint32_t tmem_base_s;
  __shared__ uint64_t bar, bar2, bar3[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // fp16 ones
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_init(&bar2, 1); for (int i = 0; i < 8; ++i) mbar_init(&bar3[i], 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (warp == 0) {
    const uint32_t idesc = idesc_f16_f32(128, N);
    const uint32_t blo = smem_desc_lo(smem_u32(smem));
    unsigned long long t0 = clock64();
    if (mode == 3 || mode == 4) {    // 12 main MMAs (N) followed by 3 projection-shaped MMAs (N = 64; mode 4: bf16 idesc)
      const uint32_t idesc_p = (mode == 4) ? idesc_bf16_f32(128, 64) : idesc_f16_f32(128, 64);
      for (int g = 0; g < groups; ++g) {
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < 3; ++k)
            mma_f16_ts(tb + 256, tb + 448, smem_desc_from_lo(blo + (k & 3) * 2), idesc_p, k != 0);
        }
        __syncwarp();
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < PER; ++k)
            mma_f16_ts(tb, tb + 456, smem_desc_from_lo(blo + (k & 3) * 2), idesc, 1);
        }
        __syncwarp();
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
    } else if (mode == 1) {
      if (lane == 0) {
        int slot = 0;
        for (int g = 0; g < groups; ++g) {
#pragma unroll
          for (int k = 0; k < PER; ++k)
            mma_f16_ts(tb, tb + 448 + slot * 8, smem_desc_from_lo(blo + (k & 3) * 2), idesc, 1);
          if (++slot == 4) slot = 0;
        }
        tc_commit(&bar);
      }
      __syncwarp();
    } else {
      int slot = 0;
      for (int g = 0; g < groups; ++g) {
        const int s = (mode == 2) ? __shfl_sync(0xffffffffu, slot, 0) : slot;
        if (mode == 6) mbar_wait(smem_u32(&bar2), 1);          // ready: fresh barrier, parity 1 passes
        if (mode == 7) {                                       // same with the non-blocking test_wait
          uint32_t ok;
          do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar2)), "r"(1) : "memory");
          } while (!ok);
        }
        if (mode == 8) {                                       // one lane polls, the rest follow through __syncwarp
          if (lane == 0) mbar_wait(smem_u32(&bar2), 1);
          __syncwarp();
        }
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < PER; ++k)
            mma_f16_ts(tb, tb + 448 + s * 8, smem_desc_from_lo(blo + (k & 3) * 2), idesc, 1);
          if (mode >= 5) tc_commit(&bar3[g & 7]);
        }
        __syncwarp();
        if (++slot == 4) slot = 0;
      }
      if (elect_one()) tc_commit(&bar);
      __syncwarp();
    }
    mbar_wait(&bar, 0);
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

template <int PER>
void run(int N, int mode, unsigned long long* d_c) {
  const int groups = 3072 / PER;
  cudaFuncSetAttribute(probe<PER>, cudaFuncAttributeMaxDynamicSharedMemorySize, 40960);
  probe<PER><<<148, 64, 40960>>>(N, mode, groups, d_c);
  cudaDeviceSynchronize();
  probe<PER><<<148, 64, 40960>>>(N, mode, groups, d_c);
  cudaError_t e = cudaDeviceSynchronize();
  unsigned long long c[148];
  cudaMemcpy(c, d_c, sizeof(c), cudaMemcpyDeviceToHost);
  printf("N=%3d mode=%d MMAs/region=%2d: %.1f cycles/MMA (floor %d), %.0f cycles/region  (%s)\n", N, mode, PER,
         (double)c[0] / (groups * PER), N / 2, (double)c[0] / groups, cudaGetErrorString(e));
}

int main() {
  unsigned long long* d_c;
  cudaMalloc(&d_c, 8 * 1024);
  for (int N : {16})
    for (int mode : {5, 6, 7, 8}) {
      run<1>(N, mode, d_c);
      run<12>(N, mode, d_c);
    }
  return 0;
}
