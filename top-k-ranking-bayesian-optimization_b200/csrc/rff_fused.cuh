// K1 — fused random-Fourier-feature evaluation, tcgen05 contraction with the posterior weight samples and
// per-sample (value, index) argmax.  Replaces fourier_features.py:18-21 + ":162-167" (n_init -> N) and the
// Phi @ theta of acquisitions/dts.py:82-85 for a basis shared by all S samples.
//
// Work item = (128-candidate tile, chunk of SC samples).  Per item the CTA accumulates
//     acc[128 x SC] = sum over 64-feature slabs of  Phi_slab[128 x 64] * ThetaScaled_slab[SC x 64]^T
// with Phi and Theta each split into fp16 hi + lo parts and three tcgen05.mma (hi*hi + hi*lo + lo*hi) per
// K=16 step, fp32 accumulation in TMEM ("fp32 emulation": ~2^-22 relative per product).
//
// Warp roles (448 threads, one CTA per SM, persistent over work items):
//   warps 0-3   epilogue: TMEM accumulator -> per-sample warp max (CREDUX) -> packed atomicMax, or store F
//   warps 4-11  feature producers (two groups of four warps; group g takes slabs g, g+2, ...): thread = one
//               candidate row; cos(x.w_j + b_j) for 64 features, hi/lo fp16 split, tcgen05.st into the TMEM
//               A-operand ring.  Phi never touches shared or global memory.
//   warp 12     TMA: Theta hi/lo slab tiles (128B-swizzled, K-major) + the W|b slab into the smem ring
//   warp 13     TMEM allocation + single-thread MMA issue + tcgen05.commit
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace tkbo {

constexpr int kTileM = 128;          // candidates per tile (TMEM lanes)
constexpr int kSlabK = 64;           // features per slab (= one 128-byte swizzle row of fp16)
constexpr int kUmmaK = 16;           // K per tcgen05.mma for 16-bit inputs
constexpr int kNumThreads = 448;
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 6;

struct RffParams {
  const float* X;            // [N, d] candidates
  const float* Wpack;        // [Dpad, WS] rows = (w_0..w_{DT-1}, b, 0...) per feature
  const float* unscale;      // [Salloc] 1 / (power-of-two scale folded into Theta)
  unsigned long long* packed;  // [Salloc] running (orderable value << 32 | ~row), zero-initialised
  unsigned int* done_counter;  // last-CTA detection
  float* out_val;            // [S]
  long long* out_idx;        // [S]
  long long* out_key;        // [S] or null: signed-orderable (value, ~global index) key instead of out_val/out_idx
  float* out_F;              // [S, ldf] (eval mode)
  long long ldf;
  long long idx_offset;
  long long N;
  int d;                     // true input dimension (row stride of X)
  int S;                     // true number of samples
  int SC;                    // samples per chunk (multiple of 32, <= 256)
  int n_chunks;
  int n_slabs;               // Dpad / 64
  int n_mtiles;
  int stages;                // ring depth (smem B/W ring and TMEM A ring)
  int acc_bufs;              // 1 or 2 accumulator buffers of SC columns
};

template <int DT> struct WStride { static constexpr int value = (DT == 1) ? 2 : ((DT + 1 + 3) / 4) * 4; };

// bytes one stage receives through its mbarrier, and the 1024-aligned stride between stages
__host__ __device__ constexpr int stage_tx_bytes(int SC, int WS) { return 2 * SC * 128 + kSlabK * WS * 4; }
__host__ __device__ constexpr int stage_stride_bytes(int SC, int WS) {
  return 2 * SC * 128 + ((kSlabK * WS * 4 + 1023) / 1024) * 1024;
}

// One candidate row, 64 features of one slab -> 32 packed fp16x2 hi words + 32 lo words.
template <int DT>
__device__ __forceinline__ void produce_slab(const float (&x)[DT], const float* __restrict__ wslab,
                                             uint32_t (&hi)[32], uint32_t (&lo)[32]) {
  constexpr int WS = WStride<DT>::value;
#pragma unroll
  for (int jj = 0; jj < 32; ++jj) {
    float c[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float* wr = wslab + (2 * jj + e) * WS;
      float w[WS];
      if constexpr (WS == 2) {
        const float2 v = *reinterpret_cast<const float2*>(wr);
        w[0] = v.x; w[1] = v.y;
      } else {
#pragma unroll
        for (int q = 0; q < WS / 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(wr + 4 * q);
          w[4 * q + 0] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
      }
      float t = w[DT];                      // phase b_j
#pragma unroll
      for (int i = 0; i < DT; ++i) t = fmaf(x[i], w[i], t);
      c[e] = __cosf(t);
    }
    const __half2 h = __floats2half2_rn(c[0], c[1]);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn(c[0] - back.x, c[1] - back.y);
    hi[jj] = *reinterpret_cast<const uint32_t*>(&h);
    lo[jj] = *reinterpret_cast<const uint32_t*>(&l);
  }
}

template <int DT, int MODE /* 0 = argmax, 1 = store F */>
__global__ void __launch_bounds__(kNumThreads, 1)
rff_fused_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                 const RffParams p) {
  constexpr int WS = WStride<DT>::value;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzled tiles; plain pointer arithmetic keeps the shared address space
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  __shared__ uint64_t bar_b_full[kMaxStages];   // TMA landed Theta+W slab
  __shared__ uint64_t bar_a_full[kMaxStages];   // producers wrote the Phi slab into TMEM
  __shared__ uint64_t bar_empty[kMaxStages];    // MMAs reading stage finished (frees smem + TMEM stage)
  __shared__ uint64_t bar_acc_full[2];
  __shared__ uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int is_last_cta;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int SC = p.SC;
  const int sbytes = stage_stride_bytes(SC, WS);
  const int txbytes = stage_tx_bytes(SC, WS);
  const int n_items = p.n_mtiles * p.n_chunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bar_b_full[s], 1);
      mbar_init(&bar_a_full[s], 128);
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_acc_full[a], 1);
      mbar_init(&bar_acc_empty[a], 128);
    }
    fence_mbar_init();
  }
  if (warp == 13) tmem_alloc(&tmem_base_smem, kTmemCols);
  if (warp == 12 && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t a_col0 = kTmemCols - p.stages * kSlabK;   // TMEM A ring occupies the top columns

  if (warp == 12) {
    // ------------------------------------------------------------------ TMA producer (warp-converged, one elected issuer)
    int st = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item % p.n_chunks;
      for (int slab = 0; slab < p.n_slabs; ++slab) {
        mbar_wait(&bar_empty[st], ph ^ 1);
        if (elect_one()) {
          uint8_t* sb = smem + st * sbytes;
          mbar_arrive_expect_tx(&bar_b_full[st], txbytes);
          tma_load_2d(sb, &tm_hi, &bar_b_full[st], slab * kSlabK, chunk * SC);
          tma_load_2d(sb + SC * 128, &tm_lo, &bar_b_full[st], slab * kSlabK, chunk * SC);
          bulk_load_1d(sb + 2 * SC * 128, p.Wpack + static_cast<size_t>(slab) * kSlabK * WS, kSlabK * WS * 4,
                       &bar_b_full[st]);
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 13) {
    // ------------------------------------------------------------------ MMA issuer (warp-converged, one elected issuer)
    const uint32_t idesc = idesc_f16_f32(kTileM, SC);
    // descriptor low words advance by plain adds: (addr >> 4) never carries out of its 14-bit field
    const uint32_t desc_lo0 = smem_desc_lo(smem_u32(smem));
    const uint32_t desc_stage = static_cast<uint32_t>(sbytes) >> 4;
    const uint32_t desc_lopart = static_cast<uint32_t>(SC * 128) >> 4;     // Theta-lo tile follows the hi tile
    int st = 0, ab = 0;
    uint32_t ph = 0, aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      mbar_wait(&bar_acc_empty[ab], aph ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * SC;
      for (int slab = 0; slab < p.n_slabs; ++slab) {
        mbar_wait(&bar_b_full[st], ph);
        mbar_wait(&bar_a_full[st], ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = desc_lo0 + st * desc_stage;
          const uint32_t b_lo = b_hi + desc_lopart;
          const uint32_t a_hi = tmem_base + a_col0 + st * kSlabK;
          const uint32_t a_lo = a_hi + kSlabK / 2;
#pragma unroll
          for (int k = 0; k < kSlabK / kUmmaK; ++k) {
            const uint32_t koff = k * ((kUmmaK * 2) >> 4);               // 32 bytes of K per step, in 16-byte units
            mma_f16_ts(d_tmem, a_hi + k * (kUmmaK / 2), smem_desc_from_lo(b_hi + koff), idesc, (slab | k) != 0);
            mma_f16_ts(d_tmem, a_hi + k * (kUmmaK / 2), smem_desc_from_lo(b_lo + koff), idesc, 1);
            mma_f16_ts(d_tmem, a_lo + k * (kUmmaK / 2), smem_desc_from_lo(b_hi + koff), idesc, 1);
          }
          tc_commit(&bar_empty[st]);
          if (slab == p.n_slabs - 1) tc_commit(&bar_acc_full[ab]);
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
      if (++ab == p.acc_bufs) { ab = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ feature producers
    // group g owns the global slab iterations it = g, g + 2, g + 4, ... (across work items), so its ring
    // position always advances by two
    const int group = (warp - 4) >> 2;               // 0 or 1
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int r = q * 32 + lane;                     // row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int st = group % p.stages;
    uint32_t ph = (group / p.stages) & 1;
    int slab = group;
    int item = blockIdx.x;
    while (slab >= p.n_slabs && item < n_items) { slab -= p.n_slabs; item += gridDim.x; }
    while (item < n_items) {
      const int mtile = item / p.n_chunks;
      long long row = static_cast<long long>(mtile) * kTileM + r;
      if (row >= p.N) row = p.N - 1;                 // clamp: tail rows are masked in the epilogue
      float x[DT];
#pragma unroll
      for (int i = 0; i < DT; ++i) x[i] = (i < p.d) ? __ldg(p.X + row * p.d + i) : 0.f;
      for (; slab < p.n_slabs; slab += 2) {
        mbar_wait(&bar_b_full[st], ph);              // W slab present; also implies the TMEM stage is free
        const float* wslab = reinterpret_cast<const float*>(smem + st * sbytes + 2 * SC * 128);
        uint32_t hi[32], lo[32];
        produce_slab<DT>(x, wslab, hi, lo);
        const uint32_t a_hi = tmem_base + lane_base + a_col0 + st * kSlabK;
        tmem_st_x32(a_hi, hi);
        tmem_st_x32(a_hi + kSlabK / 2, lo);
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bar_a_full[st]);
        st += 2;
        if (st >= p.stages) { st -= p.stages; ph ^= 1; }
      }
      slab -= p.n_slabs;                             // 0 or 1: first slab of this group in the next item
      item += gridDim.x;
      while (slab >= p.n_slabs && item < n_items) { slab -= p.n_slabs; item += gridDim.x; }   // n_slabs == 1
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp;                              // warps 0..3 <-> TMEM lanes 32q..32q+31
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int ab = 0;
    uint32_t aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int mtile = item / p.n_chunks;
      const int chunk = item % p.n_chunks;
      const long long row0 = static_cast<long long>(mtile) * kTileM + q * 32;
      const long long row = row0 + lane;
      const bool partial = (row0 + 32 > p.N);
      const bool row_ok = row < p.N;
      const int s_base = chunk * SC;

      unsigned long long bestpk[8];
      float bestv[8];
      if constexpr (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          bestpk[j] = 0ull;
          bestv[j] = -INFINITY;
          if (j * 32 < SC) {
            bestpk[j] = __ldcg(p.packed + s_base + j * 32 + lane);
            if (bestpk[j] != 0ull) bestv[j] = f32_from_orderable(static_cast<uint32_t>(bestpk[j] >> 32));
          }
        }
      }
      mbar_wait(&bar_acc_full[ab], aph);
      tc_fence_after();
      const uint32_t acc = tmem_base + lane_base + ab * SC;
#pragma unroll 1
      for (int j = 0; j < SC / 32; ++j) {
        uint32_t v[32];
        tmem_ld_x32(acc + j * 32, v);
        tmem_wait_ld();
        if constexpr (MODE == 0) {
          float mine = -INFINITY;
          if (!partial) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float m = warp_max_f32(__uint_as_float(v[c]));
              mine = (lane == c) ? m : mine;
            }
          } else {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const float m = warp_max_f32(row_ok ? __uint_as_float(v[c]) : -INFINITY);
              mine = (lane == c) ? m : mine;
            }
          }
          // bestv/bestpk are indexed by j dynamically: keep them in a small switch-free form
          float bv = -INFINITY;
          unsigned long long bp = 0ull;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (jj == j) { bv = bestv[jj]; bp = bestpk[jj]; }
          const int s = s_base + j * 32 + lane;
          const bool flag = (s < p.S) && (row0 < p.N) && (mine >= bv);
          unsigned mask = __ballot_sync(0xffffffffu, flag);
          while (mask) {
            const int c = __ffs(mask) - 1;
            mask &= mask - 1;
            const float mc = __shfl_sync(0xffffffffu, mine, c);
            float vv = __uint_as_float(tmem_ld_x1(acc + j * 32 + c));
            tmem_wait_ld();
            if (!row_ok) vv = -INFINITY;
            const unsigned hit = __ballot_sync(0xffffffffu, vv == mc);
            const int l = __ffs(hit) - 1;
            if (lane == c && hit != 0u) {
              const unsigned long long pk = pack_val_row(mc + 0.0f, static_cast<uint32_t>(row0 + l));
              if (pk > bp) {
                atomicMax(p.packed + s, pk);
                bp = pk;
                bv = mc;
              }
            }
          }
#pragma unroll
          for (int jj = 0; jj < 8; ++jj)
            if (jj == j) { bestv[jj] = bv; bestpk[jj] = bp; }
        } else {
          if (row_ok) {
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const int s = s_base + j * 32 + c;
              if (s < p.S) p.out_F[static_cast<long long>(s) * p.ldf + row] = __uint_as_float(v[c]) * __ldg(p.unscale + s);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_acc_empty[ab]);
      if (++ab == p.acc_bufs) { ab = 0; aph ^= 1; }
    }
  }

  // ---------------------------------------------------------------------- teardown + final reduce
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 13) tmem_dealloc(tmem_base, kTmemCols);
  if constexpr (MODE == 0) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(p.done_counter, 1u);
      is_last_cta = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last_cta) {
      __threadfence();
      for (int s = threadIdx.x; s < p.S; s += blockDim.x) {
        const unsigned long long pk = __ldcg(p.packed + s);
        const bool none = (pk == 0ull);           // only NaNs seen for this sample
        const float v = none ? __int_as_float(0x7fc00000)
                             : f32_from_orderable(static_cast<uint32_t>(pk >> 32)) * p.unscale[s];
        const long long gi = none ? -1 : static_cast<long long>(0xFFFFFFFFu - static_cast<uint32_t>(pk)) + p.idx_offset;
        if (p.out_key != nullptr) {
          p.out_key[s] = none ? kEmptyKey : argmax_key(v, static_cast<uint32_t>(gi));
        } else {
          p.out_val[s] = v;
          p.out_idx[s] = gi;
        }
      }
      if (threadIdx.x == 0) *p.done_counter = 0u;
    }
  }
}

}  // namespace tkbo
