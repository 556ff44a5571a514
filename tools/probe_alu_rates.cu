// Issue-rate probe for the instructions K2's inner loop is made of: MUFU.RCP, MUFU.EX2, FFMA, FFMA2, FMUL2, FADD, F2I.
// Prints warp-instructions per clock per SM sub-partition (4 per SM) with 8 warps per sub-partition and 8 independent
// chains per thread.      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_alu_rates probe_alu_rates.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, long long* cyc, int iters, float seed) {
  float v[8];
  float2 w[8];
  for (int j = 0; j < 8; ++j) { v[j] = seed + threadIdx.x * 1e-3f + j; w[j] = make_float2(v[j], v[j] + 0.5f); }
  const float2 m2 = make_float2(1.0000001f, 0.9999999f), a2 = make_float2(1e-9f, -1e-9f);
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[j]) : "f"(m2.x), "f"(a2.x));
      if (OP == 3) w[j] = __ffma2_rn(w[j], m2, a2);
      if (OP == 4) w[j] = __fmul2_rn(w[j], m2);
      if (OP == 5) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[j]) : "f"(a2.x));
      if (OP == 6) { unsigned u; asm volatile("cvt.rni.u32.f32 %0, %1;" : "=r"(u) : "f"(v[j])); v[j] = __uint_as_float(u | 0x3f800000u); }
      if (OP == 7) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 8) asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(v[j]));
      if (OP == 9) { unsigned u; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[j]), "f"(w[j].x)); v[j] = __uint_as_float((u & 0x007fffffu) | 0x3f800000u); }
      if (OP == 10) { unsigned short h; asm volatile("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(h) : "f"(v[j]), "f"(w[j].x)); v[j] = __uint_as_float(h | 0x3f800000u); }
      if (OP == 11) { asm volatile("{.reg .b16 a, m; mov.b32 {a, m}, %1; fma.rn.f32.f16 %0, a, m, %0;}" : "+f"(v[j]) : "r"(0x3c003c00u)); }
      if (OP == 13) { unsigned u; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(v[j]), "f"(w[j].x)); v[j] = __uint_as_float(u); }
      if (OP == 14) { unsigned u; asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(v[j])); asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(w[j].y), "f"(w[j].x)); w[j].x = __uint_as_float(u); }
      if (OP == 15) { asm volatile("cos.approx.ftz.f32 %0, %0;" : "+f"(v[j])); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(w[j].x) : "f"(m2.x), "f"(a2.x)); }
      if (OP == 12) { asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[j])); v[j] += 1.0f; }
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
  for (int j = 0; j < 8; ++j) s += v[j] + w[j].x + w[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name) {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  k<OP><<<148, 1024>>>(out, cyc, iters, 1.5f);
  k<OP><<<148, 1024>>>(out, cyc, iters, 1.5f);
  cudaDeviceSynchronize();
  long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
  const double winst = 8.0 * iters * 8;          // per sub-partition: 8 warps x iters x 8 chains
  printf("%-10s %.3f warp-inst/clk/sub-partition  (%.1f lanes/clk/SM)  err=%s\n", name, winst / c, winst / c * 32 * 4, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<0>("MUFU.RCP"); run<1>("MUFU.EX2"); run<7>("MUFU.LG2"); run<8>("MUFU.COS"); run<2>("FFMA"); run<3>("FFMA2"); run<4>("FMUL2"); run<5>("FADD"); run<6>("F2I"); run<9>("F2FP.F16x2"); run<10>("F2FP.E4M3x2"); run<11>("FMA.f32.f16"); run<12>("RCP+FADD"); run<13>("F2FP only"); run<14>("COS+F2FP"); run<15>("COS+FFMA");
  return 0;
}
