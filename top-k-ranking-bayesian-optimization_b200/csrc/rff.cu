// Host side of the shared-basis RFF sampler (operand preparation, tiling choice, tensor maps, launch)
// plus the small float64 kernels for the reference-exact per-sample-basis shapes.
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <math.h>

#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "rff_fused.cuh"
#include "rff_launch.cuh"

namespace tkbo {

constexpr int kHalvesDefault = 2;       // 1: the start of an item overlaps the read-out of half 1; 2: also mirror it at the end
constexpr bool kPairDefault = true;     // CTA pairs (cta_group::2) wherever they apply: 4-18 % faster on B200; TKBO_PAIR=0 switches them off

// ------------------------------------------------------------------------------------------------
// operand preparation (once per BO iteration; not the hot path)
// ------------------------------------------------------------------------------------------------
// Theta element (j, s) from either layout.
__device__ __forceinline__ double theta_at(const double* T, int transposed, int D, int S, int j, int s) {
  return transposed ? T[static_cast<size_t>(s) * D + j] : T[static_cast<size_t>(j) * S + s];
}

// bf16(v) with round-to-nearest-even, as raw bits and as the double it represents.
__device__ __forceinline__ uint16_t bf16_rn_bits(double v) {
  const float f = static_cast<float>(v);
  uint32_t u = __float_as_uint(f);
  if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);      // inf / nan
  u += 0x7FFFu + ((u >> 16) & 1u);
  return static_cast<uint16_t>(u >> 16);
}
__device__ __forceinline__ double bf16_bits_value(uint16_t h) {
  return static_cast<double>(__uint_as_float(static_cast<uint32_t>(h) << 16));
}

// All of basis_set in one launch: blocks [0, Salloc) each scale and split one sample's weights (max reduction, then the
// row's fp16 hi + fp16 lo / e4m3 correction entries), the remaining blocks write 128 rows of Wproj each.
//
// e_s is chosen so that max_j |sqrt(2/D) Theta[j,s]| 2^e_s lies in [2^top, 2^(top+1)); unscale[s] = 2^(-e_s - post) is what
// the epilogue multiplies the accumulator by (post = the power of two the feature operand carries).
// top = 9, post = 0 (three-fp16 form) keeps the fp16 hi part far from overflow and the lo part (<= 2^-11 of hi) far above
// fp16's subnormal step; top = 7, post = 12 (e4m3 correction form, features scaled by 2^12) puts the entries themselves
// (< 2^8) and the 2^12-scaled lo parts (<= 2^8) into e4m3's normal range [2^-6, 448] down to 2^-14 of the column maximum.
// e4m3 correction tile (rff_fused.cuh, C8): row s = per 16 features 16 bytes e4m3(2^12 lo part) then 16 bytes
// e4m3(value), i.e. 128 bytes per 64-feature slab; hi = the fp16 tile as in the three-MMA form.
//
// Wproj[j, :] = the K-row of feature j for the projection MMA (rff_fused.cuh): per input dimension i the six
// entries  w1 w2 w1 w3 w2 w1  (w = w1 + w2 + w3 in bf16, paired with x1 x1 x2 x1 x2 x3), then b1 b2 b3 at
// k = 6 DT .. 6 DT + 2 (paired with ones), zero padding.  Padded features get all zeros: cos(0) = 1 against
// zero Theta rows.
__global__ void __launch_bounds__(256)
prep_basis_kernel(const double* __restrict__ Theta, int transposed, const double* __restrict__ W,
                  const double* __restrict__ b, int D, int d, int S, int Dpad, int Salloc, int DT, int KA, int corr8,
                  float* __restrict__ unscale, __half* __restrict__ hi, __half* __restrict__ lo,
                  uint16_t* __restrict__ Wproj) {
  if (static_cast<int>(blockIdx.x) >= Salloc) {
    const int j = (static_cast<int>(blockIdx.x) - Salloc) * 128 + static_cast<int>(threadIdx.x);
    if (threadIdx.x >= 128 || j >= Dpad) return;
    uint16_t* row = Wproj + static_cast<size_t>(j) * KA;
    for (int k = 0; k < KA; ++k) row[k] = 0;
    if (j >= D) return;
    for (int i = 0; i <= d; ++i) {
      const double w = (i < d) ? W[static_cast<size_t>(j) * d + i] : b[j];
      const uint16_t w1 = bf16_rn_bits(w);
      const double r1 = w - bf16_bits_value(w1);
      const uint16_t w2 = bf16_rn_bits(r1);
      const uint16_t w3 = bf16_rn_bits(r1 - bf16_bits_value(w2));
      if (i < d) {
        uint16_t* o = row + 6 * i;
        o[0] = w1; o[1] = w2; o[2] = w1; o[3] = w3; o[4] = w2; o[5] = w1;
      } else {
        uint16_t* o = row + 6 * DT;
        o[0] = w1; o[1] = w2; o[2] = w3;
      }
    }
    return;
  }
  const int s = blockIdx.x;
  __shared__ double red[256];
  __shared__ double scale_sh;
  double m = 0.0;
  if (s < S)
    for (int j = threadIdx.x; j < D; j += blockDim.x) m = fmax(m, fabs(theta_at(Theta, transposed, D, S, j, s)));
  red[threadIdx.x] = m;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]);
    __syncthreads();
  }
  const double root = sqrt(2.0 / D);
  if (threadIdx.x == 0) {
    const double amax = red[0] * root;
    int e = 0;
    if (amax > 0.0 && isfinite(amax)) e = (corr8 ? 7 : 9) - ilogb(amax);
    e = max(-100, min(100, e));
    const float u = static_cast<float>(ldexp(1.0, -e - (corr8 ? 12 : 0)));
    unscale[s] = u;
    scale_sh = 1.0 / (static_cast<double>(u) * (corr8 ? 4096.0 : 1.0));
  }
  __syncthreads();
  const double sc = scale_sh;
  double l1 = 0.0;                                         // sum_j |hi operand| of this sample (two-pass argmax bound)
  for (int j = threadIdx.x; j < Dpad; j += blockDim.x) {
    double v = 0.0;
    if (s < S && j < D) v = theta_at(Theta, transposed, D, S, j, s) * root * sc;
    const __half h = __double2half(v);
    l1 += fabs(static_cast<double>(__half2float(h)));
    const size_t i = static_cast<size_t>(s) * Dpad + j;
    hi[i] = h;
    if (corr8) {
      const float l = static_cast<float>(v - static_cast<double>(__half2float(h)));
      uint8_t* row = reinterpret_cast<uint8_t*>(lo) + static_cast<size_t>(s) * 2 * Dpad + (j >> 4) * 32 + (j & 15);
      row[0] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(l * 4096.f, __NV_SATFINITE, __NV_E4M3));
      row[16] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(static_cast<float>(v), __NV_SATFINITE, __NV_E4M3));
    } else {
      lo[i] = __double2half(v - static_cast<double>(__half2float(h)));
    }
  }
  // eps2[s] (stored behind the Salloc unscale factors) = 2 eps_s in accumulator units, eps_s >= |full - hi*hi-only value|:
  // the two correction terms are at most (2^-11 + 2^-12) amp sum|v| (fp16 half-ulps of both operands, |Phi operand| <= amp
  // = 4096 in the fast form, 1 otherwise; the e4m3 corrections of the fast form approximate the same two terms), rounded
  // up to 2^-9 (the fast form's e4m3-rounded corrections reach 2^-10.3: a margin of 2.4, which costs a handful of extra items
// in pass 2); every accumulator update truncates by < 2^-23 of a partial sum that is <= amp sum|v|, and the two passes
  // make Dpad / 16 and at most 3 Dpad / 16 updates.
  __syncthreads();
  red[threadIdx.x] = l1;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double amp = corr8 ? 4096.0 : 1.0;
    const double eps = amp * red[0] * (ldexp(1.0, -9) + (Dpad / 4.0) * ldexp(1.0, -23));
    unscale[Salloc + s] = static_cast<float>(2.02 * eps);
  }
}

__global__ void merge_argmax_kernel(const float* __restrict__ vals, const long long* __restrict__ idx, int G, int S,
                                    float* __restrict__ out_val, long long* __restrict__ out_idx) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float bv = vals[s];
  long long bi = idx[s];
  for (int g = 1; g < G; ++g) {
    const float v = vals[static_cast<size_t>(g) * S + s];
    const long long i = idx[static_cast<size_t>(g) * S + s];
    if (i < 0) continue;
    if (bi < 0 || v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
  }
  out_val[s] = bv;
  out_idx[s] = bi;
}

// ------------------------------------------------------------------------------------------------
// per-sample-basis float64 kernels (reference-exact shapes: count x n_init points, count bases)
// ------------------------------------------------------------------------------------------------
// One warp per point (c, i): f = sqrt(2/D) sum_j cos(x.w_cj + b_cj) theta_cj          [ff.py:20-21,162]
__global__ void rff_eval_batched_kernel(const double* __restrict__ X, const double* __restrict__ W,
                                        const double* __restrict__ b, const double* __restrict__ theta, int count,
                                        int n, int D, int d, double* __restrict__ out_f) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= count * n) return;
  const int c = gw / n;
  const double* x = X + static_cast<size_t>(gw) * d;
  const double* Wc = W + static_cast<size_t>(c) * D * d;
  double acc = 0.0;
  for (int j = lane; j < D; j += 32) {
    double t = b[static_cast<size_t>(c) * D + j];
    for (int k = 0; k < d; ++k) t += Wc[static_cast<size_t>(j) * d + k] * x[k];
    acc += cos(t) * theta[static_cast<size_t>(c) * D + j];
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out_f[gw] = acc * sqrt(2.0 / D);
}

// phi[c, i, j] = sqrt(2/D) cos(x_ci . w_cj + b_cj)                                       [ff.py:18-21]
__global__ void rff_features_batched_kernel(const double* __restrict__ X, const double* __restrict__ W,
                                            const double* __restrict__ b, int count, int n, int D, int d,
                                            double* __restrict__ phi) {
  const size_t total = static_cast<size_t>(count) * n * D;
  const double root = sqrt(2.0 / D);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(i % D);
    const size_t ci = i / D;                       // c * n + row
    const int c = static_cast<int>(ci / n);
    double t = b[static_cast<size_t>(c) * D + j];
    for (int k = 0; k < d; ++k) t += W[(static_cast<size_t>(c) * D + j) * d + k] * X[ci * d + k];
    phi[i] = root * cos(t);
  }
}

// argmax over axis 1 of f [count, n], lowest index on ties                              [ff.py:164-165]
__global__ void argmax_rows_f64_kernel(const double* __restrict__ f, int count, int n, long long* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= count) return;
  double bv = f[static_cast<size_t>(c) * n];
  int bi = 0;
  for (int i = 1; i < n; ++i) {
    const double v = f[static_cast<size_t>(c) * n + i];
    if (v > bv) { bv = v; bi = i; }
  }
  out[c] = bi;
}

// ff.py:133-157 on the device, without a grid-wide barrier.  The update of a start point only needs that point's own
// gradient (loss = -mean over all points, so dloss/dx_p = -(1 / total) df_p/dx_p); only the early stop (ff.py:154) couples
// the points, through the global mean loss.  So the loop runs in CHUNKS of steps: one block of 128 threads per start point
// carries its point through the whole chunk (x in registers, the D features split over the threads, block reduction of f
// and df/dx per step) and records its f at every step; a reduction kernel forms the global loss per step and finds the
// first step s >= 1 with |loss_s - loss_{s-1}| / |loss_{s-1}| < rel_tol.  If there is one, the chunk is re-run from its
// starting points for exactly s steps (everything is deterministic, so the re-run reproduces the same iterates); otherwise
// the next chunk starts from the chunk's end points.  Against the cooperative one-barrier-per-step kernel of round 1 this
// has 4x the warps per point and no barrier: c1 (count 20 x n_init 50, D = 1000) 43 -> 13 us per step (the float64
// sincos of 10^6 features per step is then the bound: ~60 DFMA each on a pipe of 64 lanes per clock and SM).
constexpr int kAscentThreads = 128;
constexpr int kAscentMaxD = 16;

__global__ void __launch_bounds__(kAscentThreads)
rff_ascent_chunk_kernel(const double* __restrict__ Xin, double* __restrict__ Xout, const double* __restrict__ W,
                        const double* __restrict__ b, const double* __restrict__ theta, int n, int D, int d, int total,
                        double min_val, double max_val, int steps, double lr, double eps,
                        double* __restrict__ f_point /* [steps + 1][total] */) {
  __shared__ double red[kAscentThreads / 32][kAscentMaxD + 1];
  const int pt = blockIdx.x;
  const int c = pt / n;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const double* Wc = W + static_cast<size_t>(c) * D * d;
  const double* bc = b + static_cast<size_t>(c) * D;
  const double* tc = theta + static_cast<size_t>(c) * D;
  const double root = sqrt(2.0 / D);
  const double gcoef = root / static_cast<double>(total);
  double x[kAscentMaxD];
#pragma unroll
  for (int k = 0; k < kAscentMaxD; ++k) x[k] = (k < d) ? Xin[static_cast<size_t>(pt) * d + k] : 0.0;
  for (int step = 0; step <= steps; ++step) {
    // f and df/dx at the current x: this thread's share of the features
    double f = 0.0;
    double g[kAscentMaxD];
#pragma unroll
    for (int k = 0; k < kAscentMaxD; ++k) g[k] = 0.0;
    for (int j = threadIdx.x; j < D; j += kAscentThreads) {
      double t = bc[j];
#pragma unroll
      for (int k = 0; k < kAscentMaxD; ++k)
        if (k < d) t += Wc[static_cast<size_t>(j) * d + k] * x[k];
      double sn, cs;
      sincos(t, &sn, &cs);
      const double th = tc[j];
      f += cs * th;
      const double ts = th * sn;
#pragma unroll
      for (int k = 0; k < kAscentMaxD; ++k)
        if (k < d) g[k] += ts * Wc[static_cast<size_t>(j) * d + k];
    }
    for (int o = 16; o > 0; o >>= 1) {
      f += __shfl_xor_sync(0xffffffffu, f, o);
#pragma unroll
      for (int k = 0; k < kAscentMaxD; ++k)
        if (k < d) g[k] += __shfl_xor_sync(0xffffffffu, g[k], o);
    }
    if (lane == 0) {
      red[wib][kAscentMaxD] = f;
#pragma unroll
      for (int k = 0; k < kAscentMaxD; ++k)
        if (k < d) red[wib][k] = g[k];
    }
    __syncthreads();
    // every thread sums the warps' partials in the same order: identical f, g in the whole block
    f = 0.0;
#pragma unroll
    for (int w = 0; w < kAscentThreads / 32; ++w) f += red[w][kAscentMaxD];
#pragma unroll
    for (int k = 0; k < kAscentMaxD; ++k) {
      if (k < d) {
        double gk = 0.0;
#pragma unroll
        for (int w = 0; w < kAscentThreads / 32; ++w) gk += red[w][k];
        g[k] = gk;
      }
    }
    __syncthreads();                                  // red is rewritten by the next step
    if (threadIdx.x == 0) f_point[static_cast<size_t>(step) * total + pt] = f * root;
    if (step < steps) {
      // Keras RMSprop(rho = 0): x <- clip(x - lr g / (sqrt(g^2) + eps)), g = dloss/dx = +gcoef sum_j theta_j sin(.) w_j
#pragma unroll
      for (int k = 0; k < kAscentMaxD; ++k) {
        if (k < d) {
          const double gk = g[k] * gcoef;
          x[k] = fmin(fmax(x[k] - lr * gk / (sqrt(gk * gk) + eps), min_val), max_val);
        }
      }
    }
  }
  if (static_cast<int>(threadIdx.x) < d) {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < kAscentMaxD; ++k)
      if (k == static_cast<int>(threadIdx.x)) v = x[k];
    Xout[static_cast<size_t>(pt) * d + threadIdx.x] = v;
  }
}

// loss[s] = -(1 / total) sum_p f_point[s][p], s = 0 .. steps: one block per step, fixed summation order.
__global__ void __launch_bounds__(256)
ascent_loss_kernel(const double* __restrict__ f_point, int total, double* __restrict__ loss) {
  __shared__ double red[256];
  const double* row = f_point + static_cast<size_t>(blockIdx.x) * total;
  double s = 0.0;
  for (int i = threadIdx.x; i < total; i += 256) s += row[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) loss[blockIdx.x] = -red[0] / static_cast<double>(total);
}

// *stop = first s in 1 .. steps with |loss[s] - loss[s-1]| / |loss[s-1]| < rel_tol (ff.py:154), or -1.
__global__ void ascent_stop_kernel(const double* __restrict__ loss, int steps, double rel_tol, int* __restrict__ stop) {
  int found = -1;
  for (int s = 1; s <= steps; ++s) {
    if (fabs((loss[s] - loss[s - 1]) / loss[s - 1]) < rel_tol) { found = s; break; }
  }
  *stop = found;
}

// ------------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------------
static int pick_dt(int d) {
  const int supported[] = {1, 2, 3, 4, 6, 8, 12, 16};
  for (int v : supported)
    if (d <= v) return v;
  return -1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [Salloc, Dpad] fp16 (or, bytes = true, the [Salloc, 2 Dpad] e4m3 correction tile); box = SC rows x 128 bytes.
static int make_theta_map(tkbo_handle_t h, CUtensorMap* map, void* base, int Salloc, int Dpad, int box_rows, bool bytes = false) {
  auto fn = reinterpret_cast<EncodeTiledFn>(h->encode_tiled);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(bytes ? 2 * Dpad : Dpad), static_cast<cuuint64_t>(Salloc)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(Dpad) * sizeof(__half)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(bytes ? 2 * kSlabK : kSlabK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  CUresult r = fn(map, bytes ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return TKBO_ERR_CUDA;
  }
  return TKBO_OK;
}

// (W | b) projection rows: [Dpad, KA] bf16, box = one slab (64 features) x one 64-wide (or narrower) K block,
// swizzle span = the box row.
static int make_wproj_map(tkbo_handle_t h, CUtensorMap* map, uint16_t* base, int Dpad, int DT, int box_rows = kSlabK) {
  auto fn = reinterpret_cast<EncodeTiledFn>(h->encode_tiled);
  const int KA = proj_k_alloc(DT);
  const int row_bytes = proj_row_bytes(DT);
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(KA), static_cast<cuuint64_t>(Dpad)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(KA) * sizeof(uint16_t)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(row_bytes / 2), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estride[2] = {1, 1};
  const CUtensorMapSwizzle sw = row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, gdim, gstride, box, estride,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (W) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return TKBO_ERR_CUDA;
  }
  return TKBO_OK;
}

static void free_basis(tkbo_basis_t bs) {
  cudaFree(bs->Wproj); cudaFree(bs->theta_hi); cudaFree(bs->theta_lo); cudaFree(bs->unscale); cudaFree(bs->packed);
  delete bs;
}

int basis_create(tkbo_handle_t h, tkbo_basis_t* out, int D, int d, int S, int precision) {
  TKBO_REQUIRE(h && out, "null handle/out");
  TKBO_REQUIRE(D >= 1 && d >= 1 && S >= 1, "D, d, S must be >= 1");
  const int DT = pick_dt(d);
  if (DT < 0) {
    set_error("input dimension d > 16 is not supported by the fused kernel");
    return TKBO_ERR_UNSUPPORTED;
  }
  auto* bs = new tkbo_basis_s();
  bs->D = D; bs->d = d; bs->S = S; bs->DT = DT;
  bs->Dpad = ((D + kSlabK - 1) / kSlabK) * kSlabK;
  const int S32 = ((S + 31) / 32) * 32;
  const int smem_limit = h->max_smem_optin - 2048;      // static barriers + 1024-byte alignment slack
  const int xa_cols = proj_k_alloc(DT) / 2;
  // Sample chunk: as wide as possible (every chunk re-evaluates the features of its tile), narrowed until three
  // smem stages fit.  TMEM budget: acc_bufs * SC accumulator columns + xa_bufs * xa_cols split-candidate columns +
  // ring * 64; prefer double-buffered accumulators, then double-buffered candidate rows; the ring needs >= 3 slots
  // (projection one slab ahead of the main MMAs plus one being consumed).
  bs->ring = 0;
  for (int nc = (S32 + 255) / 256; nc <= S32 / 32 && bs->ring == 0; ++nc) {
    const int SC = ((((S32 + nc - 1) / nc) + 31) / 32) * 32;
    const int stages = std::min(kMaxStages, smem_limit / stage_stride_bytes(SC, DT));
    if (stages < 3) continue;
    for (int ab = 2; ab >= 1 && bs->ring == 0; --ab)
      for (int xb = 2; xb >= 1 && bs->ring == 0; --xb) {
        const int ring = std::min(kMaxRing, (kTmemCols - ab * SC - xb * xa_cols) / kSlabK);
        if (ring >= 3) {
          bs->ring = ring; bs->acc_bufs = ab; bs->xa_bufs = xb; bs->stages = stages;
          bs->SC = SC; bs->n_chunks = (S32 + SC - 1) / SC;
        }
      }
  }
  if (bs->ring == 0) {
    delete bs;
    set_error("no feasible pipeline configuration for this (S, d)");
    return TKBO_ERR_UNSUPPORTED;
  }
  bs->Salloc = bs->n_chunks * bs->SC;
  const int stride = stage_stride_bytes(bs->SC, DT);
  if (const char* env = getenv("TKBO_STAGES")) {
    const int t = atoi(env);
    if (t >= 3 && t <= bs->stages) bs->stages = t;
  }
  if (const char* env = getenv("TKBO_RING")) {
    const int t = atoi(env);
    if (t >= 3 && t <= bs->ring) bs->ring = t;
  }
  // two projection-issuing warps alternate slabs, which needs an even smem ring; the ring slots must not outnumber
  // the smem stages (a slot is known free through its stage's b_empty barrier)
  if (bs->stages % 2 == 1 && bs->stages >= 5) bs->stages -= 1;
  bs->proj_warps = (bs->stages % 2 == 0) ? 2 : 1;
  bs->ring = std::min(bs->ring, bs->stages);
  bs->npg = bs->ring >= 5 ? 3 : 2;
  // Small sample chunks are bound by the SFU (cos), not the tensor pipe: a fourth producer group pays off if a sixth ring
  // slot is available, which needs the split candidate rows out of TMEM (two smem buffers) and a deep smem ring.
  {
    const int ring6 = (kTmemCols - 2 * bs->SC) / kSlabK;
    const int st6 = std::min(kMaxStages, (smem_limit - 2 * proj_x_bytes(DT)) / stride) & ~1;
    const bool off = getenv("TKBO_NPG") != nullptr && atoi(getenv("TKBO_NPG")) != 4;
    if (bs->acc_bufs == 2 && ring6 >= 6 && st6 >= 6 && !off) {
      bs->npg = 4; bs->ring = std::min(kMaxRing, ring6); bs->stages = st6; bs->proj_warps = 2; bs->xa_bufs = 2;
    }
  }       // producer groups that can be busy at once: ring slots ahead of the main MMAs
  if (const char* env = getenv("TKBO_NPG")) {
    const int t = atoi(env);
    if ((t == 2 || t == 3) && bs->npg != 4) bs->npg = t;
  }
  bs->smem_bytes = bs->stages * stride + 1024 + (bs->npg == 4 ? 2 * proj_x_bytes(DT) : 0);
  // Slab-sharing producers (both groups convert every slab, half the per-slab latency) paid off only while a conversion
  // took ~9 instructions per feature; with 4-6 the fixed per-slab cost per warp outweighs it (c3 shard 6.8 vs 6.3 ms), so
  // the mode is off unless asked for.
  bs->coop = 0;
  if (const char* env = getenv("TKBO_COOP")) bs->coop = ((bs->npg == 2 || bs->npg == 4) && atoi(env) != 0) ? 1 : 0;
  // The widest sample chunks (SC > 192) leave room for only 3 combined stages, and a stage is then held from its TMA issue
  // through the projection and the feature producers until its main MMAs retire (~4500 cycles for ~1000 cycles of use).
  // Split mode gives the small (W | b) tiles their own ring: the Theta stages are held for TMA latency + main MMAs only,
  // and the projections / producers run ahead as far as the TMEM ring allows (c3 shard 7.2 -> 6.3 ms).  With 4 or more
  // combined stages the extra b_full wait of the main warps costs more than the decoupling brings (measured at
  // SC = 128 .. 192), so those keep the combined ring.
  bs->wstages = 0;
  if (bs->npg != 4 && bs->stages <= 3 && getenv("TKBO_NOSPLIT") == nullptr) {
    const int tb = 2 * bs->SC * 128, wb = proj_w_bytes(DT);
    for (int nt = std::min(kMaxStages, smem_limit / tb); nt >= 3; --nt) {
      const int nw = std::min(kMaxWStages, (smem_limit - nt * tb) / wb) & ~1;
      if (nw >= 4 || (nt == 3 && nw >= 2)) {
        bs->stages = nt; bs->wstages = nw; bs->proj_warps = 2;
        bs->ring = std::min(bs->ring, nt);
        bs->smem_bytes = nt * tb + nw * wb + 1024;
        break;
      }
    }
  }
  // CTA pairs (rff_fused.cuh, PAIR): two CTAs of a cluster share every tcgen05.mma (M = 256), each fetching half of the B
  // operands.  Needs the split candidate rows in TMEM (not the four-group form) and an even SM count; a stage is half as
  // large, so the combined ring is deep enough without the separate (W | b) ring.  TKBO_PAIR=0/1 overrides the default.
  bs->pair = 0;
  bs->halves = 0;
  bs->prune = 1;                         // 0: single pass, 1: two passes when there are >= 8 items per CTA, 2: always
  if (const char* env = getenv("TKBO_PRUNE")) bs->prune = std::max(0, std::min(2, atoi(env)));
  // two slabs ahead measured best (c3 shard, unthrottled launches: 5.80 / 5.76 / 5.72 / 5.72 ms for 4 / 1 / 2 / 3): the slabs
  // issued ahead pin their ring slots until their half-1 MMAs retire, and the producers need the other slots to keep going
  bs->lead = 2;
  if (const char* env = getenv("TKBO_LEAD")) bs->lead = std::max(0, atoi(env));
  // sample chunks of <= 128 columns are issued whole by one main MMA warp (RffParams::nsplit)
  bs->nsplit = bs->SC > 128 ? 1 : 0;
  if (const char* env = getenv("TKBO_NSPLIT")) bs->nsplit = atoi(env) != 0 ? 1 : 0;
  {
    const char* env = getenv("TKBO_PAIR");
    const bool want = env != nullptr ? atoi(env) != 0 : kPairDefault;
    if (want && bs->npg != 4 && bs->SC >= 64 && h->sm_count % 2 == 0) {
      // the halved stages leave room for the split candidate rows in shared memory (two buffers), which frees their TMEM
      // columns for one more ring slot: the cross-CTA hand-offs add latency to the projection -> producers -> main MMA
      // loop, and that loop has to fit in ring x slab time
      const int stride_p = pair_stage_stride_bytes(bs->SC, DT);
      int st = std::min(kMaxStages, (smem_limit - 2 * proj_x_bytes(DT)) / stride_p);
      const int ring = std::min(kMaxRing, (kTmemCols - bs->acc_bufs * bs->SC) / kSlabK);
      if (st >= 3 && ring >= 3) {
        bs->pair = 1; bs->stages = st; bs->wstages = 0; bs->proj_warps = 1; bs->coop = 0; bs->xa_bufs = 2;
        bs->ring = std::min(ring, st);
        // three producer groups once the ring has a slot for each; at the widest chunk (one accumulator, ring of 4) two
        // groups measured 1 % faster
        bs->npg = (bs->ring >= 5 || (bs->ring == 4 && bs->SC <= 192)) ? 3 : 2;
        if (const char* e2 = getenv("TKBO_NPG")) { const int t = atoi(e2); if (t == 2 || t == 3) bs->npg = t; }
        if (const char* e3 = getenv("TKBO_COOP")) bs->coop = (bs->npg == 2 && atoi(e3) != 0) ? 1 : 0;
        bs->smem_bytes = st * stride_p + 2 * proj_x_bytes(DT) + 1024;
        // single accumulator buffer: retire it in two column halves (rff_fused.cuh, RffParams::halves)
        const char* eh = getenv("TKBO_HALVES");
        bs->halves = (bs->acc_bufs == 1 && bs->SC % 64 == 0) ? (eh != nullptr ? atoi(eh) : kHalvesDefault) : 0;   // 0 / 1 / 2
      }
    }
  }
  // Precision of the contraction (include/tkbo.h).  AUTO is the fp32-emulating three-MMA form: the e4m3 correction form
  // (two thirds of the tensor work, ~3e-5 of max|f| on well-posed weights) is only safe when ||theta'|| / max|f| is small,
  // which the library cannot know -- callers that can bound it (fourier_features.SharedBasis) ask for FAST explicitly.
  bs->corr8 = precision == TKBO_PRECISION_FAST;
  cudaError_t e = cudaSuccess;
  e = e ? e : cudaMalloc(&bs->Wproj, static_cast<size_t>(bs->Dpad) * proj_k_alloc(DT) * sizeof(uint16_t));
  e = e ? e : cudaMalloc(&bs->theta_hi, static_cast<size_t>(bs->Salloc) * bs->Dpad * sizeof(__half));
  e = e ? e : cudaMalloc(&bs->theta_lo, static_cast<size_t>(bs->Salloc) * bs->Dpad * sizeof(__half));
  e = e ? e : cudaMalloc(&bs->unscale, static_cast<size_t>(2 * bs->Salloc) * sizeof(float));   // + eps2 (two-pass argmax)
  // + one word behind the keys: the last-CTA counter of this basis' launches (launches on different bases may
  // overlap on different streams; launches on the same basis must be stream-ordered)
  e = e ? e : cudaMalloc(&bs->packed, static_cast<size_t>(bs->Salloc + 1) * sizeof(unsigned long long));
  e = e ? e : cudaMemset(bs->packed, 0, static_cast<size_t>(bs->Salloc + 1) * sizeof(unsigned long long));
  if (e != cudaSuccess) {
    set_error(std::string("cudaMalloc (basis): ") + cudaGetErrorString(e));
    free_basis(bs);
    return TKBO_ERR_NOMEM;
  }
  const int th_rows = bs->pair ? (bs->halves ? bs->SC / 4 : bs->SC / 2) : bs->SC;      // Theta rows per TMA box
  int rc = make_theta_map(h, &bs->tm_hi, bs->theta_hi, bs->Salloc, bs->Dpad, th_rows);
  if (rc == TKBO_OK) rc = make_theta_map(h, &bs->tm_lo, bs->theta_lo, bs->Salloc, bs->Dpad, th_rows, bs->corr8 != 0);
  if (rc == TKBO_OK) rc = make_wproj_map(h, &bs->tm_w, bs->Wproj, bs->Dpad, DT, bs->pair ? kSlabK / 2 : kSlabK);
  if (rc != TKBO_OK) {
    free_basis(bs);
    return rc;
  }
  *out = bs;
  return TKBO_OK;
}

int basis_destroy(tkbo_handle_t, tkbo_basis_t bs) {
  if (!bs) return TKBO_OK;
  free_basis(bs);
  return TKBO_OK;
}

int basis_set(tkbo_handle_t h, tkbo_basis_t bs, const double* W, const double* b, const double* Theta,
              int transposed, cudaStream_t stream) {
  TKBO_REQUIRE(h && bs && W && b && Theta, "null pointer");
  const int wblocks = (bs->Dpad + 127) / 128;
  prep_basis_kernel<<<bs->Salloc + wblocks, 256, 0, stream>>>(Theta, transposed, W, b, bs->D, bs->d, bs->S, bs->Dpad,
                                                            bs->Salloc, bs->DT, proj_k_alloc(bs->DT), bs->corr8,
                                                            bs->unscale, bs->theta_hi, bs->theta_lo, bs->Wproj);
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  bs->is_set = true;
  return TKBO_OK;
}

template <int MODE>
static int dispatch_fused(tkbo_handle_t h, tkbo_basis_t bs, const RffParams& p, int grid, cudaStream_t stream) {
  switch (bs->DT) {
    case 1: return fused_launch_dt<1>(MODE, h, bs, p, grid, stream);
    case 2: return fused_launch_dt<2>(MODE, h, bs, p, grid, stream);
    case 3: return fused_launch_dt<3>(MODE, h, bs, p, grid, stream);
    case 4: return fused_launch_dt<4>(MODE, h, bs, p, grid, stream);
    case 6: return fused_launch_dt<6>(MODE, h, bs, p, grid, stream);
    case 8: return fused_launch_dt<8>(MODE, h, bs, p, grid, stream);
    case 12: return fused_launch_dt<12>(MODE, h, bs, p, grid, stream);
    case 16: return fused_launch_dt<16>(MODE, h, bs, p, grid, stream);
  }
  set_error("internal: unsupported DT");
  return TKBO_ERR_UNSUPPORTED;
}

// CTAs of a launch over `mt` candidate tiles: one per SM and item, or (pairs) two per pair of tiles and chunk, always even
static int fused_grid(tkbo_handle_t h, tkbo_basis_t bs, int64_t mt) {
  if (bs->pair) return static_cast<int>(std::min<int64_t>(h->sm_count & ~1, 2 * ((mt + 1) / 2) * bs->n_chunks));
  return static_cast<int>(std::min<int64_t>(h->sm_count, mt * bs->n_chunks));
}

static int fill_params(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, RffParams* p, int* grid) {
  TKBO_REQUIRE(h && bs && X, "null pointer");
  TKBO_REQUIRE(bs->is_set, "basis has no data: call tkbo_rff_basis_set first");
  TKBO_REQUIRE(N >= 1 && N < (int64_t(1) << 32) - 256, "N must be in [1, 2^32 - 256)");
  p->X = X; p->unscale = bs->unscale; p->packed = bs->packed;
  p->done_counter = reinterpret_cast<unsigned int*>(bs->packed + bs->Salloc);
  p->out_val = nullptr; p->out_idx = nullptr; p->out_key = nullptr; p->out_F = nullptr; p->ldf = 0; p->idx_offset = 0;
  p->duel_n = 0; p->duel_j0 = 0; p->duel_nj = 0; p->score_fix = nullptr;
  p->peer_world = 0; p->peer_rank = 0; p->peer_epoch = 0; p->ready_flag = nullptr; p->rows_per_chunk = 0;
  p->peer_merge = 0; p->peer_status = nullptr;
  p->approx = 0; p->eps2 = bs->unscale + bs->Salloc; p->rg_max = nullptr; p->lb_out = nullptr; p->item_list = nullptr; p->list_count = nullptr;
  for (int r = 0; r < 8; ++r) p->peer_buf[r] = nullptr;
  p->N = N; p->d = bs->d; p->S = bs->S; p->SC = bs->SC; p->n_chunks = bs->n_chunks;
  p->n_slabs = bs->Dpad / kSlabK;
  const int64_t mt = (N + kTileM - 1) / kTileM;
  TKBO_REQUIRE(mt * bs->n_chunks < (int64_t(1) << 31), "too many work items");
  p->n_mtiles = static_cast<int>(mt);
  p->stages = bs->stages; p->ring = bs->ring; p->acc_bufs = bs->acc_bufs; p->xa_bufs = bs->xa_bufs;
  p->proj_warps = bs->proj_warps; p->wstages = bs->wstages; p->coop = bs->coop;
  p->debug = 0;
  if (const char* env = getenv("TKBO_DEBUG")) p->debug = atoi(env);
  p->trace = h->trace; p->trace_cap = h->trace_cap;
  p->pair = bs->pair; p->halves = bs->halves; p->nsplit = bs->nsplit; p->lead = bs->lead;
  *grid = fused_grid(h, bs, mt);
  return TKBO_OK;
}

// MODE 0 launch, in two passes when the problem is large enough to pay for a second launch (RffParams::approx): pass 1 over
// every item with the hi*hi MMAs only, pass 2 in full precision over the items pass 1 could not exclude.  Results are the
// same bits as the single pass (every recorded key comes from the full-precision pass; excluded items are provably below
// the maximum).  TKBO_PRUNE=0 switches it off.
// Item list of pass 2 of the two-pass argmax (RffParams::approx): item (tile or tile pair, sample chunk) is kept if one of
// its row groups has an approximate maximum that reaches the sample's final lower bound.  One block per item, thread = column.
__global__ void __launch_bounds__(256)
mark_items_kernel(const float* __restrict__ rg_max, const float* __restrict__ lb, long long N, int S, int SC, int n_chunks,
                  int pair, int n_items, int* __restrict__ item_list, int* __restrict__ list_count) {
  const int item = blockIdx.x;
  if (item >= n_items) return;
  const int chunk = item % n_chunks;
  const long long first_tile = static_cast<long long>(item / n_chunks) * (pair ? 2 : 1);
  const long long n_groups = (N + 31) / 32;
  const long long ld = static_cast<long long>(SC) * n_chunks;
  bool keep = false;
  for (int c = threadIdx.x; c < SC && chunk * SC + c < S; c += blockDim.x) {       // (padded samples have no bound)
    const float bound = lb[chunk * SC + c];
    for (int g = 0; g < (pair ? 8 : 4); ++g) {
      const long long grp = first_tile * 4 + g;
      if (grp >= n_groups) break;
      keep |= !(rg_max[grp * ld + chunk * SC + c] < bound);          // (NaN: cannot be excluded)
    }
  }
  if (__syncthreads_or(keep) && threadIdx.x == 0) item_list[atomicAdd(list_count, 1)] = item;
}


static size_t align256(size_t b) { return (b + 255) / 256 * 256; }
static int launch_argmax(tkbo_handle_t h, tkbo_basis_t bs, const RffParams& p, int grid, cudaStream_t stream) {
  const int64_t mt = p.n_mtiles;
  const int64_t n_items = (bs->pair ? (mt + 1) / 2 : mt) * bs->n_chunks;
  // automatic: only the tensor-bound forms (sample chunks >= 128) with enough items per CTA to pay for the extra launches;
  // at c2 (chunk 64, bound by the feature map, which pass 1 evaluates in full) two passes cost 0.43 against 0.37 ms
  if (bs->prune == 0 || bs->npg == 4 || (bs->prune == 1 && (n_items < 8ll * grid || bs->SC < 128))) {
    if (h->two_pass_host) { h->two_pass_host[0] = -1; h->two_pass_host[1] = -1; }
    return dispatch_fused<0>(h, bs, p, grid, stream);
  }
  const size_t n_groups = static_cast<size_t>((p.N + 31) / 32);
  const size_t o_lb = align256(sizeof(float) * n_groups * bs->Salloc);                    // after rg_max
  const size_t o_list = o_lb + align256(sizeof(float) * bs->Salloc);
  const size_t o_cnt = o_list + align256(sizeof(int) * static_cast<size_t>(n_items));
  unsigned char* scratch = nullptr;
  if (cudaMallocAsync(&scratch, o_cnt + 256, stream) != cudaSuccess) {
    // no room for the row-group maxima (N / 32 x S floats): the single pass needs no scratch and gives the same result
    cudaGetLastError();
    if (h->two_pass_host) { h->two_pass_host[0] = -1; h->two_pass_host[1] = -1; }
    return dispatch_fused<0>(h, bs, p, grid, stream);
  }
  cudaError_t e = cudaMemsetAsync(scratch + o_cnt, 0, 256, stream);
  int rc = TKBO_OK;
  if (e != cudaSuccess) { set_error(std::string("cudaMemsetAsync: ") + cudaGetErrorString(e)); rc = TKBO_ERR_CUDA; }
  RffParams a = p;
  a.approx = 1;
  a.rg_max = reinterpret_cast<float*>(scratch);
  a.lb_out = reinterpret_cast<float*>(scratch + o_lb);
  a.peer_world = 0; a.peer_merge = 0; a.out_key = nullptr;            // pass 1 reports nothing
  if (rc == TKBO_OK) rc = dispatch_fused<0>(h, bs, a, grid, stream);
  int* list = reinterpret_cast<int*>(scratch + o_list);
  int* count = reinterpret_cast<int*>(scratch + o_cnt);
  if (rc == TKBO_OK) {
    mark_items_kernel<<<static_cast<unsigned>(n_items), 256, 0, stream>>>(a.rg_max, a.lb_out, p.N, bs->S, bs->SC, bs->n_chunks,
                                                                       bs->pair, static_cast<int>(n_items), list, count);
    h->launches += 1;
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(std::string("mark_items_kernel: ") + cudaGetErrorString(e)); rc = TKBO_ERR_CUDA; }
  }
  if (rc == TKBO_OK && h->two_pass_host) {
    h->two_pass_host[1] = static_cast<int>(n_items);
    cudaMemcpyAsync(h->two_pass_host, count, sizeof(int), cudaMemcpyDeviceToHost, stream);      // diagnostics only
  }
  RffParams c = p;
  c.item_list = list;
  c.list_count = count;
  c.ready_flag = nullptr;                                               // every chunk has landed: pass 1 has read them all
  if (rc == TKBO_OK) rc = dispatch_fused<0>(h, bs, c, grid, stream);
  e = cudaFreeAsync(scratch, stream);
  if (e != cudaSuccess && rc == TKBO_OK) { set_error(std::string("cudaFreeAsync: ") + cudaGetErrorString(e)); rc = TKBO_ERR_CUDA; }
  return rc;
}

int rff_argmax(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, int64_t idx_offset, float* out_val,
               int64_t* out_idx, cudaStream_t stream) {
  RffParams p{};
  int grid = 0;
  int rc = fill_params(h, bs, X, N, &p, &grid);
  if (rc != TKBO_OK) return rc;
  TKBO_REQUIRE(out_val && out_idx, "null output");
  p.out_val = out_val;
  p.out_idx = reinterpret_cast<long long*>(out_idx);
  p.idx_offset = idx_offset;
  return launch_argmax(h, bs, p, grid, stream);
}

// X is being filled by H2D copies on another stream, `rows_per_chunk` (multiple of 128) rows at a time; ready_flag[0]
// counts the chunks that have landed (see RffParams::ready_flag).  With peer_bufs != nullptr the launch also delivers its
// keys to every rank and merges in its last CTA (out_val / out_idx are then the GLOBAL results).
int rff_argmax_streamed(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, int64_t idx_offset, int64_t rows_per_chunk,
                        unsigned int* ready_flag, void* const* peer_bufs, int world, int rank, uint64_t epoch,
                        float* out_val, int64_t* out_idx, cudaStream_t stream) {
  RffParams p{};
  int grid = 0;
  int rc = fill_params(h, bs, X, N, &p, &grid);
  if (rc != TKBO_OK) return rc;
  TKBO_REQUIRE(out_val && out_idx, "null pointer");
  if (ready_flag != nullptr)
    TKBO_REQUIRE(rows_per_chunk >= kTileM && rows_per_chunk % kTileM == 0, "rows_per_chunk must be a multiple of 128");
  p.out_val = out_val;
  p.out_idx = reinterpret_cast<long long*>(out_idx);
  p.idx_offset = idx_offset;
  p.ready_flag = ready_flag;
  p.rows_per_chunk = rows_per_chunk;
  if (peer_bufs != nullptr) {
    TKBO_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && epoch >= 1, "bad peer arguments");
    TKBO_REQUIRE(idx_offset >= 0 && idx_offset + N <= (int64_t(1) << 32) - 1, "global index must stay below 2^32 - 1 for the key form");
    for (int r = 0; r < world; ++r) {
      TKBO_REQUIRE(peer_bufs[r] != nullptr, "null peer buffer");
      p.peer_buf[r] = static_cast<unsigned long long*>(peer_bufs[r]);
    }
    p.peer_world = world; p.peer_rank = rank; p.peer_epoch = epoch;
    p.peer_merge = 1; p.peer_status = h->peer_status_dev;
  }
  return launch_argmax(h, bs, p, grid, stream);
}

int rff_argmax_key(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, int64_t idx_offset, int64_t* out_key,
                   cudaStream_t stream) {
  RffParams p{};
  int grid = 0;
  int rc = fill_params(h, bs, X, N, &p, &grid);
  if (rc != TKBO_OK) return rc;
  TKBO_REQUIRE(out_key, "null output");
  TKBO_REQUIRE(idx_offset >= 0 && idx_offset + N <= (int64_t(1) << 32) - 1, "global index must stay below 2^32 - 1 for the key form");
  p.out_key = reinterpret_cast<long long*>(out_key);
  p.idx_offset = idx_offset;
  return launch_argmax(h, bs, p, grid, stream);
}

// ---- merge over peer memory ---------------------------------------------------------------------------------------
// Peer buffer of a rank (uint64 units): keys[2][world][S] (epoch parity, source rank, sample) then flags[8] (source rank ->
// last epoch that rank has delivered).  Rank r's fused kernel of epoch e stores its keys into slot (e & 1, r) of every
// rank's buffer and then releases flag[r] = e; the merge kernel of epoch e acquires flag[r] >= e for all r, takes the
// maximum over the `world` slots and unpacks.  Two parities suffice: a rank can only start epoch e + 2 after its merge of
// e + 1, which has seen every peer's e + 1 keys, and those were launched behind that peer's merge of e.
size_t peer_buffer_words(int world, int S) { return 2ull * world * S + 8; }

struct RffPeer {
  unsigned long long* buf[8];
  int world, rank;
  unsigned long long epoch;
};

int rff_argmax_peer(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, int64_t idx_offset, void* const* peer_bufs,
                    int world, int rank, uint64_t epoch, float* out_val, int64_t* out_idx, cudaStream_t stream) {
  RffParams p{};
  int grid = 0;
  int rc = fill_params(h, bs, X, N, &p, &grid);
  if (rc != TKBO_OK) return rc;
  TKBO_REQUIRE(peer_bufs && world >= 1 && world <= 8 && rank >= 0 && rank < world && epoch >= 1, "bad peer arguments");
  TKBO_REQUIRE(idx_offset >= 0 && idx_offset + N <= (int64_t(1) << 32) - 1, "global index must stay below 2^32 - 1 for the key form");
  for (int r = 0; r < world; ++r) {
    TKBO_REQUIRE(peer_bufs[r] != nullptr, "null peer buffer");
    p.peer_buf[r] = static_cast<unsigned long long*>(peer_bufs[r]);
  }
  p.peer_world = world; p.peer_rank = rank; p.peer_epoch = epoch;
  p.idx_offset = idx_offset;
  TKBO_REQUIRE((out_val == nullptr) == (out_idx == nullptr), "pass both out_val and out_idx (merge in the kernel) or neither");
  if (out_val != nullptr) {
    p.out_val = out_val;
    p.out_idx = reinterpret_cast<long long*>(out_idx);
    p.peer_merge = 1;
    p.peer_status = h->peer_status_dev;
  }
  return launch_argmax(h, bs, p, grid, stream);
}

// Delivery of ready-made keys (or, keys == nullptr, of "nothing seen": a rank whose shard is empty still has to release
// its flag): one block, same stores / fence / release as the fused kernel's last CTA.
__global__ void peer_push_keys_kernel(const long long* __restrict__ keys, int S, RffPeer pr) {
  const size_t slot0 = (static_cast<size_t>(pr.epoch & 1ull) * pr.world + pr.rank) * S;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const long long key = keys ? keys[s] : kEmptyKey;
    for (int r = 0; r < pr.world; ++r) pr.buf[r][slot0 + s] = static_cast<unsigned long long>(key);
  }
  __threadfence_system();
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < pr.world) {
    unsigned long long* flag = pr.buf[threadIdx.x] + 2ull * pr.world * S + pr.rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(pr.epoch) : "memory");
  }
}

int peer_push_keys(tkbo_handle_t h, const int64_t* keys, int S, void* const* peer_bufs, int world, int rank, uint64_t epoch,
                   cudaStream_t stream) {
  TKBO_REQUIRE(h && peer_bufs && S >= 1, "null pointer");
  TKBO_REQUIRE(world >= 1 && world <= 8 && rank >= 0 && rank < world && epoch >= 1, "bad peer arguments");
  RffPeer pr;
  for (int r = 0; r < 8; ++r) pr.buf[r] = r < world ? static_cast<unsigned long long*>(peer_bufs[r]) : nullptr;
  for (int r = 0; r < world; ++r) TKBO_REQUIRE(pr.buf[r] != nullptr, "null peer buffer");
  pr.world = world; pr.rank = rank; pr.epoch = epoch;
  peer_push_keys_kernel<<<1, 256, 0, stream>>>(reinterpret_cast<const long long*>(keys), S, pr);
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

// One block per 256 samples; its first `world` threads wait for the flags (bounded: ~10 s, then every index of the block
// reads -2 so that a lost peer shows up as an error instead of a hang).
__global__ void merge_peer_keys_kernel(const unsigned long long* __restrict__ buf, int world, int S, unsigned long long epoch,
                                       float* __restrict__ out_val, long long* __restrict__ out_idx,
                                       unsigned int* __restrict__ status) {
  __shared__ int timed_out;
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  if (threadIdx.x < world) {
    const unsigned long long* flag = buf + 2ull * world * S + threadIdx.x;
    const long long t0 = clock64();
    for (;;) {
      unsigned long long f;
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(flag) : "memory");
      if (f >= epoch) break;
      if (clock64() - t0 > 20000000000ll) { timed_out = 1; break; }
      __nanosleep(64);
    }
  }
  __syncthreads();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  if (timed_out) {
    out_val[s] = __int_as_float(0x7fc00000); out_idx[s] = -2;
    if (s == 0 && status != nullptr) *reinterpret_cast<volatile unsigned int*>(status) = static_cast<unsigned int>(epoch);
    return;
  }
  long long best = kEmptyKey;
  const unsigned long long* keys = buf + static_cast<size_t>(epoch & 1ull) * world * S;
  for (int r = 0; r < world; ++r) {
    unsigned long long k;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(k) : "l"(keys + static_cast<size_t>(r) * S + s) : "memory");
    best = max(best, static_cast<long long>(k));
  }
  float v;
  long long gi;
  argmax_unkey(best, &v, &gi);
  out_val[s] = v;
  out_idx[s] = gi;
}

int merge_peer_keys(tkbo_handle_t h, const void* local_buf, int world, int S, uint64_t epoch, float* out_val, int64_t* out_idx,
                    cudaStream_t stream) {
  TKBO_REQUIRE(h && local_buf && out_val && out_idx, "null pointer");
  TKBO_REQUIRE(world >= 1 && world <= 8 && S >= 1 && epoch >= 1, "bad peer arguments");
  merge_peer_keys_kernel<<<(S + 255) / 256, 256, 0, stream>>>(static_cast<const unsigned long long*>(local_buf), world, S, epoch,
                                                               out_val, reinterpret_cast<long long*>(out_idx),
                                                               h->peer_status_dev);
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

__global__ void unpack_key_kernel(const long long* __restrict__ key, int S, float* __restrict__ out_val,
                                  long long* __restrict__ out_idx) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float v;
  long long gi;
  argmax_unkey(key[s], &v, &gi);
  out_val[s] = v;
  out_idx[s] = gi;
}

// key[s] = max over g of keys[g, s] (largest value, then lowest global index), unpacked.
__global__ void merge_unpack_keys_kernel(const long long* __restrict__ keys, int G, int S, float* __restrict__ out_val,
                                         long long* __restrict__ out_idx) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  long long best = keys[s];
  for (int g = 1; g < G; ++g) best = max(best, keys[static_cast<size_t>(g) * S + s]);
  float v;
  long long gi;
  argmax_unkey(best, &v, &gi);
  out_val[s] = v;
  out_idx[s] = gi;
}

int merge_unpack_keys(tkbo_handle_t h, const int64_t* keys, int G, int S, float* out_val, int64_t* out_idx,
                      cudaStream_t stream) {
  TKBO_REQUIRE(h && keys && out_val && out_idx && G >= 1 && S >= 1, "bad merge arguments");
  merge_unpack_keys_kernel<<<(S + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const long long*>(keys), G, S, out_val,
                                                              reinterpret_cast<long long*>(out_idx));
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int unpack_argmax_key(tkbo_handle_t h, const int64_t* key, int S, float* out_val, int64_t* out_idx, cudaStream_t stream) {
  TKBO_REQUIRE(h && key && out_val && out_idx && S >= 1, "bad unpack arguments");
  unpack_key_kernel<<<(S + 127) / 128, 128, 0, stream>>>(reinterpret_cast<const long long*>(key), S, out_val,
                                                        reinterpret_cast<long long*>(out_idx));
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int rff_eval(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, float* out_F, int64_t ldf,
             cudaStream_t stream) {
  RffParams p{};
  int grid = 0;
  int rc = fill_params(h, bs, X, N, &p, &grid);
  if (rc != TKBO_OK) return rc;
  TKBO_REQUIRE(out_F && ldf >= N, "out_F null or ldf < N");
  p.out_F = out_F;
  p.ldf = ldf;
  return dispatch_fused<1>(h, bs, p, grid, stream);
}

// score[s, i] = fix[s, i] / (2^23 n), argmax over i (lowest index on ties)                 [dts.py:94-97]
__global__ void copeland_finish_kernel(const unsigned long long* __restrict__ fix, int n, float* __restrict__ score,
                                       long long* __restrict__ out_arg) {
  const int s = blockIdx.x;
  __shared__ unsigned long long bv[256];
  __shared__ int bi[256];
  unsigned long long best = 0ull;
  int idx = 0x7fffffff;
  const double inv = 1.0 / (static_cast<double>(kCopelandFixScale) * static_cast<double>(n));
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const unsigned long long v = fix[static_cast<size_t>(s) * n + i];
    if (score) score[static_cast<size_t>(s) * n + i] = static_cast<float>(static_cast<double>(v) * inv);
    if (idx == 0x7fffffff || v > best) { best = v; idx = i; }       // strided ascending: first maximum of this thread
  }
  bv[threadIdx.x] = best; bi[threadIdx.x] = idx;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const unsigned long long v = bv[threadIdx.x + o];
      const int j = bi[threadIdx.x + o];
      if (j != 0x7fffffff && (bi[threadIdx.x] == 0x7fffffff || v > bv[threadIdx.x] || (v == bv[threadIdx.x] && j < bi[threadIdx.x]))) {
        bv[threadIdx.x] = v; bi[threadIdx.x] = j;
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0 && out_arg) out_arg[s] = bi[0];
}

// Dueling-Thompson soft-Copeland sums over the ordered pairs (i, j), j in [j_begin, j_end), of n base points; the duels are
// generated on the fly and fix[s, i] += sum_j round(sigmoid(f_s([x_i, x_j])) * 2^23) (the caller zeroes fix; shards of j held by
// different ranks add up exactly with an integer all-reduce).
int rff_duel_copeland_partial(tkbo_handle_t h, tkbo_basis_t bs, const float* Xbase, int n, int j_begin, int j_end,
                              uint64_t* fix, cudaStream_t stream) {
  TKBO_REQUIRE(h && bs && Xbase && fix, "null pointer");
  TKBO_REQUIRE(bs->is_set, "basis has no data: call tkbo_rff_basis_set first");
  TKBO_REQUIRE(bs->d % 2 == 0, "duel mode needs an even input dimension d = 2 d_obj");
  TKBO_REQUIRE(n >= 1 && n <= (1 << 20), "n must be in [1, 2^20]");
  TKBO_REQUIRE(0 <= j_begin && j_begin <= j_end && j_end <= n, "need 0 <= j_begin <= j_end <= n");
  if (j_begin == j_end) return TKBO_OK;
  RffParams p{};
  const int64_t tpi = (j_end - j_begin + kTileM - 1) / kTileM;       // tiles per i: 128 consecutive j each
  const int64_t mt = static_cast<int64_t>(n) * tpi;
  TKBO_REQUIRE(mt * bs->n_chunks < (int64_t(1) << 31), "too many work items");
  p.X = Xbase; p.unscale = bs->unscale; p.packed = bs->packed;
  p.done_counter = reinterpret_cast<unsigned int*>(bs->packed + bs->Salloc);
  p.out_val = nullptr; p.out_idx = nullptr; p.out_key = nullptr; p.out_F = nullptr; p.ldf = 0; p.idx_offset = 0;
  p.N = mt * kTileM; p.d = bs->d; p.S = bs->S; p.SC = bs->SC; p.n_chunks = bs->n_chunks;
  p.n_slabs = bs->Dpad / kSlabK; p.n_mtiles = static_cast<int>(mt);
  p.stages = bs->stages; p.ring = bs->ring; p.acc_bufs = bs->acc_bufs; p.xa_bufs = bs->xa_bufs;
  p.proj_warps = bs->proj_warps; p.wstages = bs->wstages; p.coop = bs->coop;
  p.peer_world = 0; p.peer_rank = 0; p.peer_epoch = 0; p.ready_flag = nullptr; p.rows_per_chunk = 0;
  p.peer_merge = 0; p.peer_status = nullptr;
  for (int r = 0; r < 8; ++r) p.peer_buf[r] = nullptr;
  p.duel_n = n; p.duel_j0 = j_begin; p.duel_nj = j_end - j_begin;
  p.score_fix = reinterpret_cast<unsigned long long*>(fix);
  p.debug = 0;
  p.trace = nullptr; p.trace_cap = 0;
  p.pair = bs->pair; p.halves = bs->halves; p.nsplit = bs->nsplit; p.lead = bs->lead;
  const int grid = fused_grid(h, bs, mt);
  return dispatch_fused<2>(h, bs, p, grid, stream);
}

int copeland_finish(tkbo_handle_t h, const uint64_t* fix, int S, int n, float* out_score, int64_t* out_argmax,
                    cudaStream_t stream) {
  TKBO_REQUIRE(h && fix && S >= 1 && n >= 1 && (out_score || out_argmax), "bad copeland_finish arguments");
  copeland_finish_kernel<<<S, 256, 0, stream>>>(reinterpret_cast<const unsigned long long*>(fix), n, out_score,
                                                reinterpret_cast<long long*>(out_argmax));
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int rff_duel_copeland(tkbo_handle_t h, tkbo_basis_t bs, const float* Xbase, int n, float* out_score, int64_t* out_argmax,
                      cudaStream_t stream) {
  TKBO_REQUIRE(h && bs, "null pointer");
  TKBO_REQUIRE(out_score || out_argmax, "nothing to compute");
  TKBO_REQUIRE(n >= 1 && n <= (1 << 20), "n must be in [1, 2^20]");
  uint64_t* fix = nullptr;
  const size_t bytes = sizeof(uint64_t) * static_cast<size_t>(bs->S) * n;
  TKBO_CUDA_CHECK(cudaMallocAsync(&fix, bytes, stream));
  TKBO_CUDA_CHECK(cudaMemsetAsync(fix, 0, bytes, stream));
  int rc = rff_duel_copeland_partial(h, bs, Xbase, n, 0, n, fix, stream);
  if (rc == TKBO_OK) rc = copeland_finish(h, fix, bs->S, n, out_score, out_argmax, stream);
  TKBO_CUDA_CHECK(cudaFreeAsync(fix, stream));
  return rc;
}

int basis_config(tkbo_handle_t h, tkbo_basis_t bs, int64_t N, int32_t* cfg) {
  TKBO_REQUIRE(h && bs && cfg, "null pointer");
  const int64_t mt = (N + kTileM - 1) / kTileM;
  cfg[0] = bs->SC; cfg[1] = bs->n_chunks; cfg[2] = bs->Dpad / kSlabK; cfg[3] = bs->stages; cfg[4] = bs->acc_bufs;
  cfg[5] = bs->DT; cfg[6] = bs->smem_bytes;
  cfg[7] = fused_grid(h, bs, std::max<int64_t>(1, mt));
  cfg[8] = bs->ring; cfg[9] = bs->pair ? kNumThreadsPair(bs->npg) : kNumThreadsFor(bs->npg); cfg[10] = bs->xa_bufs; cfg[11] = bs->npg; cfg[12] = bs->proj_warps; cfg[13] = bs->corr8; cfg[14] = bs->wstages; cfg[15] = bs->coop + 2 * bs->pair;
  return TKBO_OK;
}

int merge_argmax(tkbo_handle_t h, const float* vals, const int64_t* idx, int G, int S, float* out_val,
                 int64_t* out_idx, cudaStream_t stream) {
  TKBO_REQUIRE(h && vals && idx && out_val && out_idx && G >= 1 && S >= 1, "bad merge arguments");
  merge_argmax_kernel<<<(S + 127) / 128, 128, 0, stream>>>(vals, reinterpret_cast<const long long*>(idx), G, S, out_val,
                                                          reinterpret_cast<long long*>(out_idx));
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int rff_eval_batched(tkbo_handle_t h, const double* X, const double* W, const double* b, const double* theta,
                     int count, int n, int D, int d, double* out_f, int64_t* out_idx, cudaStream_t stream) {
  TKBO_REQUIRE(h && X && W && b && theta, "null pointer");
  TKBO_REQUIRE(count >= 1 && n >= 1 && D >= 1 && d >= 1, "count, n, D, d must be >= 1");
  TKBO_REQUIRE(out_f || out_idx, "nothing to compute");
  double* f = out_f;
  if (!f) TKBO_CUDA_CHECK(cudaMallocAsync(&f, sizeof(double) * count * n, stream));
  const long long threads = static_cast<long long>(count) * n * 32;
  rff_eval_batched_kernel<<<static_cast<int>((threads + 255) / 256), 256, 0, stream>>>(X, W, b, theta, count, n, D, d, f);
  h->launches += 1;
  if (out_idx) {
    argmax_rows_f64_kernel<<<(count + 63) / 64, 64, 0, stream>>>(f, count, n, reinterpret_cast<long long*>(out_idx));
    h->launches += 1;
  }
  if (!out_f) TKBO_CUDA_CHECK(cudaFreeAsync(f, stream));
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int rff_features_batched(tkbo_handle_t h, const double* X, const double* W, const double* b, int count, int n, int D,
                         int d, double* out_phi, cudaStream_t stream) {
  TKBO_REQUIRE(h && X && W && b && out_phi, "null pointer");
  TKBO_REQUIRE(count >= 1 && n >= 1 && D >= 1 && d >= 1, "count, n, D, d must be >= 1");
  const size_t total = static_cast<size_t>(count) * n * D;
  const int blocks = static_cast<int>(std::min<size_t>((total + 255) / 256, static_cast<size_t>(h->sm_count) * 32));
  rff_features_batched_kernel<<<blocks, 256, 0, stream>>>(X, W, b, count, n, D, d, out_phi);
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int rff_ascent_batched(tkbo_handle_t h, double* X, const double* W, const double* b, const double* theta, int count,
                       int n, int D, int d, double min_val, double max_val, int num_steps, double lr, double eps,
                       double rel_tol, int* steps_run, cudaStream_t stream) {
  TKBO_REQUIRE(h && X && W && b && theta, "null pointer");
  TKBO_REQUIRE(count >= 1 && n >= 1 && D >= 1 && d >= 1 && d <= kAscentMaxD, "count, n, D >= 1 and 1 <= d <= 16 required");
  TKBO_REQUIRE(num_steps >= 0, "num_steps must be >= 0");
  const int total = count * n;
  int chunk = 64;       // measured at c1 with an early stop at step 856: 64 -> 13.2, 256 -> 15.3, 1024 -> 25 us per step
  if (const char* env = getenv("TKBO_ASCENT_CHUNK")) { const int t = atoi(env); if (t >= 1 && t <= 4096) chunk = t; }
  chunk = std::max(1, std::min(chunk, std::max(num_steps, 1)));
  // two point buffers (ping-pong between chunks) | f_point [chunk + 1][total] | loss [chunk + 1] | stop
  const size_t pts = static_cast<size_t>(total) * d;
  const size_t n_doubles = 2 * pts + static_cast<size_t>(chunk + 1) * total + (chunk + 1);
  double* work = nullptr;
  TKBO_CUDA_CHECK(cudaMallocAsync(&work, sizeof(double) * n_doubles + sizeof(int), stream));
  double* xa = work;
  double* xb = work + pts;
  double* f_point = xb + pts;
  double* loss = f_point + static_cast<size_t>(chunk + 1) * total;
  int* stop_dev = reinterpret_cast<int*>(loss + (chunk + 1));
  auto fail = [&](cudaError_t e, const char* what) {
    set_error(std::string(what) + ": " + cudaGetErrorString(e));
    cudaFreeAsync(work, stream);
    return TKBO_ERR_CUDA;
  };
  cudaError_t e = cudaMemcpyAsync(xa, X, sizeof(double) * pts, cudaMemcpyDeviceToDevice, stream);
  if (e != cudaSuccess) return fail(e, "cudaMemcpyAsync");
  int done = 0;
  while (done < num_steps) {
    const int c = std::min(chunk, num_steps - done);
    rff_ascent_chunk_kernel<<<total, kAscentThreads, 0, stream>>>(xa, xb, W, b, theta, n, D, d, total, min_val, max_val, c, lr,
                                                                  eps, f_point);
    ascent_loss_kernel<<<c + 1, 256, 0, stream>>>(f_point, total, loss);
    ascent_stop_kernel<<<1, 1, 0, stream>>>(loss, c, rel_tol, stop_dev);
    h->launches += 3;
    int stop = -1;
    e = cudaMemcpyAsync(&stop, stop_dev, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return fail(e, "ascent chunk");
    if (stop < 0) {
      std::swap(xa, xb);                      // the chunk's end points start the next chunk
      done += c;
    } else {
      // the early stop fell inside this chunk: the same start points, exactly `stop` steps
      rff_ascent_chunk_kernel<<<total, kAscentThreads, 0, stream>>>(xa, xb, W, b, theta, n, D, d, total, min_val, max_val, stop,
                                                                    lr, eps, f_point);
      h->launches += 1;
      std::swap(xa, xb);
      done += stop;
      break;
    }
  }
  e = cudaMemcpyAsync(X, xa, sizeof(double) * pts, cudaMemcpyDeviceToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return fail(e, "ascent");
  TKBO_CUDA_CHECK(cudaFreeAsync(work, stream));
  if (steps_run) *steps_run = done;
  return TKBO_OK;
}

}  // namespace tkbo
