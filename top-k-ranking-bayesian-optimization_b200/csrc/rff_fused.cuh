// K1 — fused random-Fourier-feature evaluation, tcgen05 contraction with the posterior weight samples and
// per-sample (value, index) argmax.  Replaces fourier_features.py:18-21 + ":162-167" (n_init -> N) and the
// Phi @ theta of acquisitions/dts.py:82-85 for a basis shared by all S samples.
//
// Work item = (group of TILES 128-candidate tiles, chunk of SC samples).  Per item and tile the CTA accumulates
//     acc[128 x SC] = sum over 64-feature slabs of  Phi_slab[128 x 64] * ThetaScaled_slab[SC x 64]^T
// with Phi and Theta each split into fp16 hi + lo parts and three tcgen05.mma (hi*hi + hi*lo + lo*hi) per
// K=16 step, fp32 accumulation in TMEM ("fp32 emulation": ~2^-22 relative per product).
//
// TILES = 2 (chosen by the host when SC <= 64, where the feature producers rather than the tensor pipe bound the
// kernel): every producer thread evaluates TWO candidate rows per W|b value it reads, which halves the
// shared-memory broadcast traffic and the L2 -> smem Theta traffic per candidate.
//
// Warp roles (32 * (6 + 4 NPG) threads, one CTA per SM, persistent over work items):
//   warps 0-3   epilogue: TMEM accumulator -> per-sample warp max (CREDUX) -> packed atomicMax, or store F
//   warps 4..   feature producers (NPG groups of four warps; group g takes global slab iterations g, g+NPG, ...):
//               thread = one candidate row per tile; cos(x.w_j + b_j) for 64 features, hi/lo fp16 split,
//               tcgen05.st into the TMEM A-operand ring.  Phi never touches shared or global memory.
//   next warp   TMA: Theta hi/lo slab tiles (128B-swizzled, K-major) + the W|b slab into the smem ring
//   last warp   TMEM allocation + one elected thread issuing tcgen05.mma / tcgen05.commit
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace tkbo {

constexpr int kTileM = 128;          // candidates per tile (TMEM lanes)
constexpr int kSlabK = 64;           // features per slab (= one 128-byte swizzle row of fp16)
constexpr int kUmmaK = 16;           // K per tcgen05.mma for 16-bit inputs
constexpr int kNumThreadsFor(int npg) { return 32 * (4 + 4 * npg + 2); }   // epilogue + producers + TMA + MMA warps
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
  int n_mtiles;              // ceil(N / 128)
  int stages;                // ring depth (smem B/W ring and TMEM A ring)
  int acc_bufs;              // 1 or 2 accumulator buffers of TILES * SC columns
  int debug;                 // timing ablations (TKBO_DEBUG): 1 = issue only the hi*hi MMA, 2 = producers skip the math
  unsigned long long* trace; // optional clock64() event log of CTA 0 ([kTraceEvents][trace_cap]), tools/trace_k1.py
  int trace_cap;
};

// Pipeline trace events (CTA 0, lane 0 of one warp per role); index = that role's own iteration counter.
enum TraceEvent : int {
  kTrTmaEmpty = 0, kTrTmaIssued = 1,
  kTrProdBFull = 2 /* + 2 * group */, kTrProdArrived = 3 /* + 2 * group */,
  kTrMmaBFull = 8, kTrMmaAFull = 9, kTrMmaCommitted = 10,
  kTrEpiAccFull = 11, kTrEpiDone = 12, kTrMmaAccEmpty = 13, kTraceEvents = 14
};
__device__ __forceinline__ void trace_event(const RffParams& p, int ev, int idx) {
  if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && idx < p.trace_cap)
    p.trace[ev * p.trace_cap + idx] = static_cast<unsigned long long>(clock64());
}

template <int DT> struct WStride { static constexpr int value = (DT == 1) ? 2 : ((DT + 1 + 3) / 4) * 4; };

// bytes one stage receives through its mbarrier, and the 1024-aligned stride between stages
__host__ __device__ constexpr int stage_tx_bytes(int SC, int WS) { return 2 * SC * 128 + kSlabK * WS * 4; }
__host__ __device__ constexpr int stage_stride_bytes(int SC, int WS) {
  return 2 * SC * 128 + ((kSlabK * WS * 4 + 1023) / 1024) * 1024;
}

// TILES candidate rows x 32 features (half a slab) -> 16 packed fp16x2 hi words + 16 lo words per row.
// One W|b read serves all TILES rows.
template <int DT, int TILES>
__device__ __forceinline__ void produce_half_slab(const float (&x)[TILES][DT], const float* __restrict__ wslab,
                                                  uint32_t (&hi)[TILES][16], uint32_t (&lo)[TILES][16]) {
  constexpr int WS = WStride<DT>::value;
#pragma unroll
  for (int jj = 0; jj < 16; ++jj) {
    float c[TILES][2];
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
#pragma unroll
      for (int t = 0; t < TILES; ++t) {
        float arg = w[DT];                    // phase b_j
#pragma unroll
        for (int i = 0; i < DT; ++i) arg = fmaf(x[t][i], w[i], arg);
        c[t][e] = __cosf(arg);
      }
    }
#pragma unroll
    for (int t = 0; t < TILES; ++t) {
      const __half2 h = __floats2half2_rn(c[t][0], c[t][1]);
      const float2 back = __half22float2(h);
      const __half2 l = __floats2half2_rn(c[t][0] - back.x, c[t][1] - back.y);
      hi[t][jj] = *reinterpret_cast<const uint32_t*>(&h);
      lo[t][jj] = *reinterpret_cast<const uint32_t*>(&l);
    }
  }
}

template <int DT, int MODE /* 0 = argmax, 1 = store F */, int TILES /* 1 or 2 */, int NPG /* producer warp groups: 2 or 3 */>
__global__ void __launch_bounds__(kNumThreadsFor(NPG), 1)
rff_fused_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                 const RffParams p) {
  constexpr int WS = WStride<DT>::value;
  constexpr int kAStage = TILES * kSlabK;            // TMEM columns per A-ring stage: per tile 32 hi + 32 lo
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
  const int n_groups = (p.n_mtiles + TILES - 1) / TILES;
  const int n_items = n_groups * p.n_chunks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bar_b_full[s], 1);
      mbar_init(&bar_a_full[s], 4);          // one arrival per producer warp of the owning group
      mbar_init(&bar_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bar_acc_full[a], 1);
      mbar_init(&bar_acc_empty[a], 4);       // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  constexpr int kTmaWarp = 4 + 4 * NPG;
  constexpr int kMmaWarp = kTmaWarp + 1;
  if (warp == kMmaWarp) tmem_alloc(&tmem_base_smem, kTmemCols);
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t a_col0 = kTmemCols - p.stages * kAStage;   // TMEM A ring occupies the top columns

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer (warp-converged, one elected issuer)
    int st = 0, it = 0;
    uint32_t ph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int chunk = item % p.n_chunks;
      for (int slab = 0; slab < p.n_slabs; ++slab, ++it) {
        mbar_wait(&bar_empty[st], ph ^ 1);
        trace_event(p, kTrTmaEmpty, it);
        if (elect_one()) {
          uint8_t* sb = smem + st * sbytes;
          if (p.debug & 4) {               // ablation: no Theta tile loads
            mbar_arrive_expect_tx(&bar_b_full[st], kSlabK * WS * 4);
          } else {
            mbar_arrive_expect_tx(&bar_b_full[st], txbytes);
            tma_load_2d(sb, &tm_hi, &bar_b_full[st], slab * kSlabK, chunk * SC);
            tma_load_2d(sb + SC * 128, &tm_lo, &bar_b_full[st], slab * kSlabK, chunk * SC);
          }
          bulk_load_1d(sb + 2 * SC * 128, p.Wpack + static_cast<size_t>(slab) * kSlabK * WS, kSlabK * WS * 4,
                       &bar_b_full[st]);
        }
        __syncwarp();
        trace_event(p, kTrTmaIssued, it);
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------------ MMA issuer (warp-converged, one elected issuer)
    const uint32_t idesc = idesc_f16_f32(kTileM, SC);
    // descriptor low words advance by plain adds: (addr >> 4) never carries out of its 14-bit field
    const uint32_t desc_lo0 = smem_desc_lo(smem_u32(smem));
    const uint32_t desc_stage = static_cast<uint32_t>(sbytes) >> 4;
    const uint32_t desc_lopart = static_cast<uint32_t>(SC * 128) >> 4;     // Theta-lo tile follows the hi tile
    int st = 0, ab = 0, it = 0, nitem = 0;
    uint32_t ph = 0, aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++nitem) {
      mbar_wait(&bar_acc_empty[ab], aph ^ 1);
      trace_event(p, kTrMmaAccEmpty, nitem);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + ab * (TILES * SC);
      for (int slab = 0; slab < p.n_slabs; ++slab, ++it) {
        mbar_wait(&bar_b_full[st], ph);
        trace_event(p, kTrMmaBFull, it);
        mbar_wait(&bar_a_full[st], ph);
        trace_event(p, kTrMmaAFull, it);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t b_hi = desc_lo0 + st * desc_stage;
          const uint32_t b_lo = b_hi + desc_lopart;
          const uint32_t a_st = tmem_base + a_col0 + st * kAStage;
#pragma unroll
          for (int k = 0; k < kSlabK / kUmmaK; ++k) {
            const uint32_t koff = k * ((kUmmaK * 2) >> 4);               // 32 bytes of K per step, in 16-byte units
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
              const uint32_t a_hi = a_st + t * kSlabK + k * (kUmmaK / 2);
              const uint32_t a_lo = a_hi + kSlabK / 2;
              const uint32_t d = d_tmem + t * SC;
              mma_f16_ts(d, a_hi, smem_desc_from_lo(b_hi + koff), idesc, (slab | k) != 0);
              if (!(p.debug & 1)) {
                mma_f16_ts(d, a_hi, smem_desc_from_lo(b_lo + koff), idesc, 1);
                mma_f16_ts(d, a_lo, smem_desc_from_lo(b_hi + koff), idesc, 1);
              }
            }
          }
          tc_commit(&bar_empty[st]);
          if (slab == p.n_slabs - 1) tc_commit(&bar_acc_full[ab]);
        }
        __syncwarp();
        trace_event(p, kTrMmaCommitted, it);
        if (++st == p.stages) { st = 0; ph ^= 1; }
      }
      if (++ab == p.acc_bufs) { ab = 0; aph ^= 1; }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------------ feature producers
    // group g owns the global slab iterations it = g, g + NPG, g + 2 NPG, ... (across work items), so its ring
    // position always advances by NPG
    const int group = (warp - 4) >> 2;               // 0 .. NPG-1
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int r = q * 32 + lane;                     // row inside the tile
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int st = group % p.stages;
    uint32_t ph = (group / p.stages) & 1;
    int slab = group;
    int item = blockIdx.x;
    int it = 0;
    while (slab >= p.n_slabs && item < n_items) { slab -= p.n_slabs; item += gridDim.x; }
    while (item < n_items) {
      const int grp = item / p.n_chunks;
      float x[TILES][DT];
#pragma unroll
      for (int t = 0; t < TILES; ++t) {
        long long row = (static_cast<long long>(grp) * TILES + t) * kTileM + r;
        if (row >= p.N) row = p.N - 1;               // clamp: tail rows are masked in the epilogue
#pragma unroll
        for (int i = 0; i < DT; ++i) x[t][i] = (i < p.d) ? __ldg(p.X + row * p.d + i) : 0.f;
      }
      for (; slab < p.n_slabs; slab += NPG, ++it) {
        mbar_wait(&bar_b_full[st], ph);              // W slab present; also implies the TMEM stage is free
        if (q == 0) trace_event(p, kTrProdBFull + 2 * group, it);
        const float* wslab = reinterpret_cast<const float*>(smem + st * sbytes + 2 * SC * 128);
        const uint32_t a_st = tmem_base + lane_base + a_col0 + st * kAStage;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t hi[TILES][16], lo[TILES][16];
          if (p.debug & 2) {
#pragma unroll
            for (int t = 0; t < TILES; ++t)
#pragma unroll
              for (int jj = 0; jj < 16; ++jj) { hi[t][jj] = 0x3c003c00u; lo[t][jj] = 0u; }
          } else {
            produce_half_slab<DT, TILES>(x, wslab + half * 32 * WS, hi, lo);
          }
          if (!(p.debug & 8)) {
#pragma unroll
            for (int t = 0; t < TILES; ++t) {
              tmem_st_x16(a_st + t * kSlabK + half * 16, hi[t]);
              tmem_st_x16(a_st + t * kSlabK + kSlabK / 2 + half * 16, lo[t]);
            }
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_a_full[st]);
        if (q == 0) trace_event(p, kTrProdArrived + 2 * group, it);
        st += NPG;
        if (st >= p.stages) { st -= p.stages; ph ^= 1; }
      }
      slab -= p.n_slabs;                             // first slab of this group in the next item
      item += gridDim.x;
      while (slab >= p.n_slabs && item < n_items) { slab -= p.n_slabs; item += gridDim.x; }   // n_slabs == 1
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const int q = warp;                              // warps 0..3 <-> TMEM lanes 32q..32q+31
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    int ab = 0, nitem = 0;
    uint32_t aph = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++nitem) {
      const int grp = item / p.n_chunks;
      const int chunk = item % p.n_chunks;
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
      if (q == 0) trace_event(p, kTrEpiAccFull, nitem);
      tc_fence_after();
#pragma unroll 1
      for (int t = 0; t < TILES; ++t) {
        const long long row0 = (static_cast<long long>(grp) * TILES + t) * kTileM + q * 32;
        if (row0 >= p.N) continue;                   // whole warp beyond the last candidate (warp-uniform)
        const long long row = row0 + lane;
        const bool partial = (row0 + 32 > p.N);
        const bool row_ok = row < p.N;
        const uint32_t acc = tmem_base + lane_base + ab * (TILES * SC) + t * SC;
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
            // bestv/bestpk are indexed by j dynamically: select without local-memory indexing
            float bv = -INFINITY;
            unsigned long long bp = 0ull;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              if (jj == j) { bv = bestv[jj]; bp = bestpk[jj]; }
            const int s = s_base + j * 32 + lane;
            const bool flag = (s < p.S) && (mine >= bv);
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_acc_empty[ab]);
      if (q == 0) trace_event(p, kTrEpiDone, nitem);
      if (++ab == p.acc_bufs) { ab = 0; aph ^= 1; }
    }
  }

  // ---------------------------------------------------------------------- teardown + final reduce
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, kTmemCols);
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
