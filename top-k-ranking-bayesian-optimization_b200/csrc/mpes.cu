// K2 — MPES Monte-Carlo ranking-likelihood mutual information, given the f-samples.
// Replaces acquisitions/rank_pes.py:38-105 (top-k ranking, Plackett-Luce likelihood over all
// k-permutations) and acquisitions/pes.py:66-123 (top-1 with eps-smoothed counts), plus the
// soft-Copeland tail of acquisitions/dts.py:94-97.
//
// Streaming form: samples are grouped by the maximiser that wins them (fstar_argmax).  For sample s and
// query q, with e_c = exp(f_c - max_c f_c), the probability of the partial ranking o (ascending
// preference, rank_pes.py:139) is   prod_{i=k-1..0} e[o_i] / (sum of e over the not-yet-ranked choices);
// this equals exp of the log-likelihood built at rank_pes.py:160-166.  Per query we accumulate
// J(z, m) = (1/S) sum_{s in group m} P_s(z) in fp32 registers, and at each group boundary fold
//     MI += J log J - J log p(m)      and      tot(z) += J(z, m)                     (fp64 logs),
// finishing with  MI -= sum_z tot(z) log tot(z), which is rank_pes.py:97-105 term by term.
#include <cuda_fp16.h>
#include <math.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "common.h"
#include "ptx.cuh"

namespace tkbo {

constexpr double kPesEps = 1e-7;   // tf.keras.backend.epsilon(), pes.py:68,87

// ------------------------------------------------------------------------------------------------
// stable counting sort of the sample indices by winning maximiser
// ------------------------------------------------------------------------------------------------
// order[group_start[m] .. group_start[m+1]) = samples with fstar_argmax == m, ascending.
// chunk_group[0 .. J]: the groups split into J contiguous chunks of roughly equal sample counts (work units of the
// all-permutations kernel; a chunk may be empty).
__global__ void group_samples_kernel(const int* __restrict__ arg, int S, int M, int* __restrict__ order,
                                     int* __restrict__ group_start, int J = 0, int* __restrict__ chunk_group = nullptr) {
  extern __shared__ int cnt[];           // [M + 1]
  for (int m = threadIdx.x; m <= M; m += blockDim.x) cnt[m] = 0;
  __syncthreads();
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int m = arg[s];
    if (m >= 0 && m < M) atomicAdd(&cnt[m + 1], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (int m = 0; m < M; ++m) cnt[m + 1] += cnt[m];
  __syncthreads();
  for (int m = threadIdx.x; m <= M; m += blockDim.x) group_start[m] = cnt[m];
  if (threadIdx.x == 0 && chunk_group != nullptr) {
    int m = 0;
    chunk_group[0] = 0;
    for (int j = 1; j < J; ++j) {
      const long long target = static_cast<long long>(cnt[M]) * j / J;
      while (m < M && cnt[m + 1] <= target) ++m;       // groups wholly below the j-th cut
      chunk_group[j] = m;
    }
    chunk_group[J] = M;
  }
  for (int m = threadIdx.x; m < M; m += blockDim.x) {
    int w = cnt[m];
    for (int s = 0; s < S; ++s)
      if (arg[s] == m) order[w++] = s;
  }
}

__global__ void argmax_rows_f32_kernel(const float* __restrict__ f, int S, int M, int* __restrict__ out) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  float bv = f[static_cast<size_t>(s) * M];
  int bi = 0;
  for (int m = 1; m < M; ++m) {
    const float v = f[static_cast<size_t>(s) * M + m];
    if (v > bv) { bv = v; bi = m; }
  }
  out[s] = bi;
}

// ------------------------------------------------------------------------------------------------
// compile-time enumeration of all K-permutations of C choices
// ------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int perm_count(int C, int K) {
  int p = 1;
  for (int i = 0; i < K; ++i) p *= (C - i);
  return p;
}
__host__ __device__ constexpr int popcount_c(unsigned m) {
  int n = 0;
  for (; m; m &= m - 1) ++n;
  return n;
}

__device__ __forceinline__ float fast_ex2(float x) {      // raw MUFU.EX2 (flushes denormals; inputs here are <= 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_rcp(float x) {      // raw MUFU.RCP, ~1 ulp
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 1 / (sum of e over the choices NOT in mask) for every mask with fewer than K members: the Plackett-Luce
// denominators depend on the *set* of already-ranked choices only, so C=5, K=3 needs 1 + 5 + 10 reciprocals for
// its 60 orderings.  Sums are taken directly over the remaining entries (no T - e_a cancellation).
template <int C, int K>
__device__ __forceinline__ void subset_reciprocals(const float (&e)[C], float (&rs)[1 << C]) {
#pragma unroll
  for (unsigned m = 0; m < (1u << C); ++m) {
    if (popcount_c(m) >= K || popcount_c(m) >= C - 1) continue;       // a lone remaining choice has probability 1
    float s = 0.f;
    bool first = true;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if ((m >> c) & 1u) continue;
      s = first ? e[c] : s + e[c];
      first = false;
    }
    rs[m] = fast_rcp(fmaxf(s, 1.17549435e-38f));          // all remaining choices flushed to zero: 0 * 2^126 = 0, not NaN
  }
}

template <int C, int K, int DEPTH>
struct PermWalk {
  // prob = probability of the choices ranked so far (best first); used = their bit mask.
  template <int P>
  static __device__ __forceinline__ void run(const float (&e)[C], const float (&rs)[1 << C], unsigned used, float prob,
                                             float (&acc)[P], int& leaf) {
    constexpr int remaining = C - DEPTH;
    const float pr = (remaining > 1) ? ((DEPTH == 0) ? rs[0] : prob * rs[used]) : prob;   // else factor e/e = 1
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if ((used >> c) & 1u) continue;
      if constexpr (DEPTH + 1 == K) {
        acc[leaf] = (remaining > 1) ? fmaf(pr, e[c], acc[leaf]) : acc[leaf] + pr;
        ++leaf;
      } else {
        const float child = (remaining > 1) ? pr * e[c] : pr;
        PermWalk<C, K, DEPTH + 1>::run(e, rs, used | (1u << c), child, acc, leaf);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// fast kernel: one thread per query, all K-permutations of C in registers.  Blocks of four independent warps
// (one per SM sub-partition), three blocks per SM, persistent over 32-query tiles: an equal warp count on every
// sub-partition matters because the loop is issue / MUFU bound.  Each warp streams its 32 queries' sample rows
// (32*C contiguous floats per sample) through a private shared-memory ring filled by 1-D bulk copies (TMA)
// kRing rows ahead, or by plain loads when the rows are not 16-byte aligned / the tile is ragged (STAGED = false).
// ------------------------------------------------------------------------------------------------
constexpr int kRing = 4;
constexpr int kMpesWarps = 4;
constexpr int kMaxChunks = 8;
__host__ __device__ constexpr int mpes_warp_smem_bytes(int C, int P) {
  return ((kRing * 32 * C * 4 + P * 32 * 4 + kRing * 8 + 127) / 128) * 128;
}

// Accurate log(a) for the group fold.  MI is a difference of entropies of size ~log P with a result of 1e-9 .. 1e-1
// (the eps-smoothed pes.I of a query that cannot move the maximiser is ~1e-9), so the per-group sums
// sum_z a_z log a_z  need ~1e-12 relative accuracy; MUFU.LG2 (2^-22) is far short and a float64 log() per value would
// cost a fifth of the kernel.  Instead: a = 2^E m, m in [1, 2); k = top 7 mantissa bits; r_k ~ 1 / (1 + (k + 1/2)/128)
// with 9 significant bits; p = fl(m r_k), e = fma(m, r_k, -p) (the exact product error), xh = p - 1 (exact, |xh| < 0.006):
//     log m = -log r_k (float64 table) + xh + [e (1 - xh) + xh^2 (-1/2 + xh/3 - xh^2/4)]
// where the bracket is below 2e-5, so float32 is enough for it; everything else is added in float64.  ~20 instructions
// per value, 4 of them on the float64 pipe; absolute error ~1e-12.
struct LogTable {
  float r[128];
  double t[128];      // -log(r[k])
};
__device__ __forceinline__ void log_table_fill(LogTable& tab, int tid) {
  if (tid < 128) {
    const float r0 = 1.f / (1.f + (static_cast<float>(tid) + 0.5f) * (1.f / 128.f));
    const float r = __uint_as_float(__float_as_uint(r0) & 0xFFFFC000u);       // 9 explicit mantissa bits
    tab.r[tid] = r;
    tab.t[tid] = -log(static_cast<double>(r));
  }
}
// log(a) for a normal positive float a
__device__ __forceinline__ double table_log(const LogTable& tab, float a) {
  const uint32_t bits = __float_as_uint(a);
  const int E = static_cast<int>(bits >> 23) - 127;
  const uint32_t k = (bits >> 16) & 127u;
  const float m = __uint_as_float((bits & 0x007FFFFFu) | 0x3F800000u);
  const float r = tab.r[k];
  const float p = __fmul_rn(m, r);
  const float e = __fmaf_rn(m, r, -p);
  const float xh = p - 1.f;
  const float q = fmaf(xh, fmaf(xh, -0.25f, 0.33333334f), -0.5f);
  const float rest = fmaf(xh * xh, q, fmaf(-e, xh, e));
  return fma(static_cast<double>(E), 0.6931471805599453, tab.t[k] + (static_cast<double>(xh) + static_cast<double>(rest)));
}

// Work unit = (tile of 32 queries, chunk j of the maximiser groups), fetched dynamically from a global counter: a
// whole tile (all S samples) is ~0.5 ms of work for one warp, far too coarse for a few thousand persistent warps.
// For its chunk a unit produces
//     A_j  = sum_{m in j} p_m sum_z c_zm log c_zm      (t1buf[j, q], float64)
//     tot_j[z] = sum_{m in j} sum_{s in m} P_s(z)       (totbuf[j, z, q], float32)
// and the LAST unit of a tile to finish (tile_done counter, threadfence + atomic) assembles
//     MI = sum_j A_j - sum_z p_z log p_z,  p_z = (sum_j tot_j[z] + M eps) / S                 (float64)
// summing j in order, so the result does not depend on which unit came last.  J == 1 keeps everything in the unit.
// Per group: a_z = acc_z + eps, c_zm = a_z cm with cm = 1 / (S p_m), hence
//     p_m sum_z c log c = (1/S) (sum_z a_z log a_z + log(cm) sum_z a_z)
// with the two sums over z in float64 (table_log above).
template <int C, int K, bool STAGED>
__global__ void __launch_bounds__(32 * kMpesWarps, 4)
mpes_all_perms_kernel(const float* __restrict__ fchi, const int* __restrict__ order, const int* __restrict__ group_start,
                      const int* __restrict__ chunk_group, int J, int S, long long Q, long long Qpad, int M, int mode,
                      int tile_begin, int tile_end, unsigned int* __restrict__ unit_counter,
                      unsigned int* __restrict__ tile_done, float* __restrict__ totbuf, double* __restrict__ t1buf,
                      double* __restrict__ out_mi) {
  constexpr int P = perm_count(C, K);
  constexpr float kLog2e = 1.4426950408889634f;
  // exponent offset: e_c = 2^((f_c - max) log2e + 120), so that choices up to 246 binades (170 nats) below the best one stay
  // normal floats (the sums stay below 2^127, every stage factor e * rs is scale free); rank_pes.py:166 renormalises per
  // stage, which this reproduces as long as the not-yet-ranked choices do not ALL flush to zero
  constexpr float kExpBias = 120.f;
  __shared__ LogTable logtab;
  // dynamic shared memory, one private slice per warp: ring | tot | mbarriers (see mpes_warp_smem_bytes)
  extern __shared__ __align__(128) unsigned char mpes_smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;                        // after the table fill, warps never synchronise with each other
  log_table_fill(logtab, threadIdx.x);
  __syncthreads();
  unsigned char* slice = mpes_smem + static_cast<size_t>(wib) * mpes_warp_smem_bytes(C, P);
  float (*ring)[32 * C] = reinterpret_cast<float (*)[32 * C]>(slice);
  float (*tot)[32] = reinterpret_cast<float (*)[32]>(slice + kRing * 32 * C * 4);      // sum over the chunk's groups of acc
  uint64_t* full = reinterpret_cast<uint64_t*>(tot + P);
  const double invS = 1.0 / static_cast<double>(S);
  const double eps = (mode == TKBO_MPES_PES) ? kPesEps : 0.0;
  const double meps = (mode == TKBO_MPES_PES) ? M * kPesEps : 0.0;
  const long long rowstride = Q * C;
  if constexpr (STAGED) {
    if (lane == 0)
      for (int i = 0; i < kRing; ++i) mbar_init(&full[i], 1);
    fence_mbar_init();
    __syncwarp();
  }
  uint32_t ring_phase = 0;                                 // bit i = parity to wait for on slot i
  const unsigned n_units = static_cast<unsigned>(tile_end - tile_begin) * static_cast<unsigned>(J);

  for (;;) {
    unsigned unit = 0;
    if (lane == 0) unit = atomicAdd(unit_counter, 1u);
    unit = __shfl_sync(0xffffffffu, unit, 0);
    if (unit >= n_units) break;
    const int tile = tile_begin + static_cast<int>(unit / J);
    const int j = static_cast<int>(unit % J);
    const int m0 = chunk_group[j], m1 = chunk_group[j + 1];
    const int i0 = group_start[m0], i1 = group_start[m1];  // this unit's range of the grouped sample order
    const long long q = static_cast<long long>(tile) * 32 + lane;
    const bool valid = q < Q;
    const float* base = fchi + (valid ? q : Q - 1) * C;                          // direct path
    const float* tile_base = fchi + static_cast<long long>(tile) * 32 * C;       // staged path
#pragma unroll
    for (int z = 0; z < P; ++z) tot[z][lane] = 0.f;
    float acc[P];
#pragma unroll
    for (int z = 0; z < P; ++z) acc[z] = 0.f;
    double A = 0.0;

    if constexpr (STAGED) {                                // prime the ring with the unit's first kRing rows
      if (lane == 0) {
        for (int r = 0; r < kRing && i0 + r < i1; ++r) {
          mbar_arrive_expect_tx(&full[r], 32 * C * 4);
          bulk_load_1d(ring[r], tile_base + static_cast<long long>(order[i0 + r]) * rowstride, 32 * C * 4, &full[r]);
        }
      }
      __syncwarp();
    }
    float fnext[C];
    if constexpr (!STAGED) {
      if (i0 < i1) {
        const float* r = base + static_cast<long long>(order[i0]) * rowstride;
#pragma unroll
        for (int c = 0; c < C; ++c) fnext[c] = __ldg(r + c);
      }
    }

    int i = i0;                                            // position in the grouped sample order
    for (int m = m0; m < m1; ++m) {
      const int hi = group_start[m + 1];
      const int cnt = hi - group_start[m];
      if (cnt == 0 && mode == TKBO_MPES_RANK) continue;    // rank_pes.py:86-93 drops empty groups
      for (; i < hi; ++i) {
        float f[C];
        if constexpr (STAGED) {
          const int slot = (i - i0) % kRing;
          mbar_wait(&full[slot], (ring_phase >> slot) & 1u);
          ring_phase ^= 1u << slot;
#pragma unroll
          for (int c = 0; c < C; ++c) f[c] = ring[slot][lane * C + c];
        } else {
#pragma unroll
          for (int c = 0; c < C; ++c) f[c] = fnext[c];
          if (i + 1 < i1) {
            const float* r = base + static_cast<long long>(order[i + 1]) * rowstride;
#pragma unroll
            for (int c = 0; c < C; ++c) fnext[c] = __ldg(r + c);
          }
        }
        float mx = f[0];
#pragma unroll
        for (int c = 1; c < C; ++c) mx = fmaxf(mx, f[c]);
        const float mxs = fmaf(mx, kLog2e, -kExpBias);
        float e[C];
#pragma unroll
        for (int c = 0; c < C; ++c) e[c] = fast_ex2(fmaf(f[c], kLog2e, -mxs));
        if constexpr (STAGED) {
          __syncwarp();                                    // every lane has consumed its slot values
          if (lane == 0 && i + kRing < i1) {
            const int slot = (i - i0) % kRing;
            mbar_arrive_expect_tx(&full[slot], 32 * C * 4);
            bulk_load_1d(ring[slot], tile_base + static_cast<long long>(order[i + kRing]) * rowstride, 32 * C * 4, &full[slot]);
          }
        }
        float rs[1 << C];
        subset_reciprocals<C, K>(e, rs);
        int leaf = 0;
        PermWalk<C, K, 0>::run(e, rs, 0u, 1.f, acc, leaf);
      }
      // fold group m:  A += p_m sum_z c_zm log c_zm  with c_zm = p(z | m) = (acc_z + eps) cm
      const double pm = (mode == TKBO_MPES_PES)
                            ? (static_cast<double>(cnt) + kPesEps) / (static_cast<double>(S) + M * kPesEps)
                            : static_cast<double>(cnt) * invS;
      const double lcm = log(invS / pm);
      double s1a = 0.0, s1b = 0.0, s0 = 0.0;              // two interleaved chains for the float64 pipe
#pragma unroll
      for (int z = 0; z < P; ++z) {
        // a = acc + eps in float64 (eps = 1e-7 is below float32 resolution of a sum of ~S/M probabilities, and pes.I of
        // an uninformative query consists of nothing but these eps terms): log a = log(fl32(a)) + (a - fl32(a)) / fl32(a)
        const double ad = static_cast<double>(acc[z]) + eps;
        const float af = static_cast<float>(ad);
        if (af >= 1e-30f) {                                // below: a log a < 7e-29, and no denormal handling needed
          double la = table_log(logtab, af);
          if (mode == TKBO_MPES_PES) la += (ad - static_cast<double>(af)) * static_cast<double>(fast_rcp(af));
          const double t = ad * la;
          if (z & 1) s1b += t; else s1a += t;
          s0 += ad;
        }
        tot[z][lane] += acc[z];
        acc[z] = 0.f;
      }
      A += invS * ((s1a + s1b) + lcm * s0);
    }
    if (J == 1) {
      // the unit saw every group: finish here, nothing goes through global memory
      double mi = A;
#pragma unroll
      for (int z = 0; z < P; ++z) {
        const double pz = (static_cast<double>(tot[z][lane]) + meps) * invS;
        if (pz > 0.0) mi -= pz * log(pz);
      }
      if (valid) out_mi[q] = mi;
      continue;
    }
#pragma unroll
    for (int z = 0; z < P; ++z) totbuf[(static_cast<size_t>(j) * P + z) * Qpad + q] = tot[z][lane];
    t1buf[static_cast<size_t>(j) * Qpad + q] = A;
    // last unit of the tile to get here assembles MI (release: partials before the counter; acquire: counter before reads)
    __threadfence();
    __syncwarp();
    unsigned prev = 0;
    if (lane == 0) prev = atomicAdd(tile_done + tile, 1u);
    prev = __shfl_sync(0xffffffffu, prev, 0);
    if (prev == static_cast<unsigned>(J - 1)) {
      __threadfence();
      double mi = 0.0;
      for (int jj = 0; jj < J; ++jj) mi += __ldcg(t1buf + static_cast<size_t>(jj) * Qpad + q);
#pragma unroll 4
      for (int z = 0; z < P; ++z) {
        double t = 0.0;
        for (int jj = 0; jj < J; ++jj) t += static_cast<double>(__ldcg(totbuf + (static_cast<size_t>(jj) * P + z) * Qpad + q));
        const double pz = (t + meps) * invS;
        if (pz > 0.0) mi -= pz * log(pz);
      }
      if (valid) out_mi[q] = mi;
      if (lane == 0) tile_done[tile] = 0u;               // ready for the next call that reuses the buffer
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic kernel: explicit permutation table (any C <= 8, k <= C, P <= 1024); one warp per query,
// lanes stride over table rows, <= 32 rows per lane in registers.
// ------------------------------------------------------------------------------------------------
constexpr int kGenMaxC = 8;
constexpr int kGenMaxP = 1024;

__global__ void __launch_bounds__(128)
mpes_table_kernel(const float* __restrict__ fchi, const int* __restrict__ order, const int* __restrict__ group_start,
                  int S, long long Q, int C, int M, const int* __restrict__ perms, int P, int k, int mode,
                  double* __restrict__ out_mi) {
  extern __shared__ unsigned char sperm[];          // [P][8] choice indices, ascending preference
  for (int i = threadIdx.x; i < P * kGenMaxC; i += blockDim.x) {
    const int p = i / kGenMaxC, j = i % kGenMaxC;
    sperm[i] = (j < k) ? static_cast<unsigned char>(perms[p * k + j]) : 0;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const long long q = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (q >= Q) return;
  const double invS = 1.0 / static_cast<double>(S);
  const double eps = (mode == TKBO_MPES_PES) ? kPesEps : 0.0;
  const float* base = fchi + q * C;
  const long long rowstride = Q * C;
  float acc[32], tot[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) { acc[i] = 0.f; tot[i] = 0.f; }
  double mi = 0.0;
  for (int m = 0; m < M; ++m) {
    const int lo = group_start[m], hi = group_start[m + 1];
    if (hi == lo && mode == TKBO_MPES_RANK) continue;
    for (int i = lo; i < hi; ++i) {
      const float* r = base + static_cast<long long>(order[i]) * rowstride;
      const float fmine = (lane < C) ? __ldg(r + lane) : -INFINITY;
      float mx = fmine;
      for (int o = 4; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));   // C <= 8
      mx = __shfl_sync(0xffffffffu, mx, 0);
      const float emine = (lane < C) ? __expf(fmine - mx + 80.f) : 0.f;   // recentred: choices down to e^-167 of the best stay normal
#pragma unroll
      for (int t = 0; t < 32; ++t) {
        const int p = lane + 32 * t;
        const bool act = p < P;
        const unsigned char* row = sperm + (act ? p : 0) * kGenMaxC;
        unsigned used = 0u;
        float prob = 1.f;
        if (t * 32 < P) {                              // warp-uniform
          for (int j = k - 1; j >= 0; --j) {
            const int c = row[j];
            float rem = 0.f;
            for (int cc = 0; cc < C; ++cc) {
              const float ev = __shfl_sync(0xffffffffu, emine, cc);
              if (!((used >> cc) & 1u)) rem += ev;
            }
            const float ec = __shfl_sync(0xffffffffu, emine, c);
            prob *= ec / fmaxf(rem, 1.17549435e-38f);
            used |= 1u << c;
          }
          if (act) acc[t] += prob;
        }
      }
    }
    const double pm = (mode == TKBO_MPES_PES)
                          ? (static_cast<double>(hi - lo) + kPesEps) / (static_cast<double>(S) + M * kPesEps)
                          : static_cast<double>(hi - lo) * invS;
    const double log_pm = log(pm);
#pragma unroll
    for (int t = 0; t < 32; ++t) {
      if (lane + 32 * t < P) {
        const double J = (static_cast<double>(acc[t]) + eps) * invS;
        if (J > 0.0) mi += J * (log(J) - log_pm);
        tot[t] += acc[t];
      }
      acc[t] = 0.f;
    }
  }
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    if (lane + 32 * t < P) {
      const double pz = (static_cast<double>(tot[t]) + M * eps) * invS;
      if (pz > 0.0) mi -= pz * log(pz);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mi += __shfl_xor_sync(0xffffffffu, mi, o);
  if (lane == 0) out_mi[q] = mi;
}

// ------------------------------------------------------------------------------------------------
// top-1 MPES with an indifference outcome (acquisitions/indiff_pes.py:15-130): C + 1 observation types,
//   P(i) = e^{f_i} / (e^{f_i} + sum_{j != i} e^{f_j + t}),   P(none) = clip(1 - sum_i P(i))        [:107-118]
// same grouped Monte-Carlo MI as rank_pes (:44-92).  One thread per query, fp32 probabilities summed per group,
// fp64 fold (C + 1 <= 9 logs per group).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
mpes_indiff_kernel(const float* __restrict__ fchi, const int* __restrict__ order, const int* __restrict__ group_start,
                   int S, long long Q, int C, int M, float tau /* e^threshold */, double* __restrict__ out_mi) {
  const long long q = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const float* base = fchi + q * C;
  const long long rowstride = Q * C;
  const double invS = 1.0 / static_cast<double>(S);
  float acc[kGenMaxC + 1], tot[kGenMaxC + 1];
#pragma unroll
  for (int z = 0; z <= kGenMaxC; ++z) { acc[z] = 0.f; tot[z] = 0.f; }
  double mi = 0.0;
  for (int m = 0; m < M; ++m) {
    const int lo = group_start[m], hi = group_start[m + 1];
    if (hi == lo) continue;                                // indiff_pes.py:68-76 drops maximisers that win no sample
    for (int i = lo; i < hi; ++i) {
      const float* r = base + static_cast<long long>(order[i]) * rowstride;
      float e[kGenMaxC];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kGenMaxC; ++c)
        if (c < C) { e[c] = __ldg(r + c); mx = fmaxf(mx, e[c]); }
      float T = 0.f;
#pragma unroll
      for (int c = 0; c < kGenMaxC; ++c)
        if (c < C) { e[c] = __expf(e[c] - mx); T += e[c]; }
      float psum = 0.f;
#pragma unroll
      for (int c = 0; c < kGenMaxC; ++c)
        if (c < C) {
          const float pc = e[c] / (e[c] + tau * (T - e[c]));
          acc[c] += pc;
          psum += pc;
        }
      acc[kGenMaxC] += fminf(fmaxf(1.f - psum, 0.f), 1.f);
    }
    const double pm = static_cast<double>(hi - lo) * invS;
    const double log_pm = log(pm);
#pragma unroll
    for (int z = 0; z <= kGenMaxC; ++z) {
      if (z < C || z == kGenMaxC) {
        const double J = static_cast<double>(acc[z]) * invS;
        if (J > 0.0) mi += J * (log(J) - log_pm);
        tot[z] += acc[z];
      }
      acc[z] = 0.f;
    }
  }
#pragma unroll
  for (int z = 0; z <= kGenMaxC; ++z) {
    if (z < C || z == kGenMaxC) {
      const double pz = static_cast<double>(tot[z]) * invS;
      if (pz > 0.0) mi -= pz * log(pz);
    }
  }
  out_mi[q] = mi;
}

// ------------------------------------------------------------------------------------------------
// DTS tail: soft-Copeland score and its argmax                               (acquisitions/dts.py:94-97)
// ------------------------------------------------------------------------------------------------
__global__ void soft_copeland_rows_kernel(const float* __restrict__ f, int n, float* __restrict__ score) {
  const int i = blockIdx.x;
  __shared__ double red[256];
  double s = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) s += 1.0 / (1.0 + exp(-static_cast<double>(f[static_cast<size_t>(i) * n + j])));
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) score[i] = static_cast<float>(red[0] / n);
}

__global__ void argmax_vec_kernel(const float* __restrict__ v, int n, long long* __restrict__ out) {
  __shared__ float bv[256];
  __shared__ int bi[256];
  float best = -INFINITY;
  int idx = 0x7fffffff;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = v[i];
    if (x > best || (x == best && i < idx)) { best = x; idx = i; }
  }
  bv[threadIdx.x] = best; bi[threadIdx.x] = idx;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      const float x = bv[threadIdx.x + o];
      const int j = bi[threadIdx.x + o];
      if (x > bv[threadIdx.x] || (x == bv[threadIdx.x] && j < bi[threadIdx.x])) { bv[threadIdx.x] = x; bi[threadIdx.x] = j; }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = bi[0];
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
// Partial results of the all-permutations kernel (one slab per sample chunk), its two unit counters and the per-tile
// "units done" counters that elect the finishing unit.
struct MpesParts {
  const int* chunk_group;   // [J + 1]
  int J;
  long long Qpad;           // queries rounded up to whole 32-query tiles
  float* totbuf;            // [J, P, Qpad]   (J > 1 only)
  double* t1buf;            // [J, Qpad]      (J > 1 only)
  unsigned int* counters;   // [2], zeroed before the launches
  unsigned int* tile_done;  // [tiles], zeroed before the launches (J > 1 only)
};

template <int C, int K>
static void launch_all_perms(const float* fchi, const int* order, const int* gs, int S, int64_t Q, int M, int mode,
                             int sm_count, const MpesParts& pt, double* out, cudaStream_t stream) {
  const int64_t tiles = (Q + 31) / 32;
  // the bulk-copy path needs 16-byte aligned rows (Q*C % 4 == 0, base aligned) and full 32-query tiles
  const bool aligned = ((Q * C) % 4 == 0) && ((reinterpret_cast<uintptr_t>(fchi) & 15) == 0);
  const int64_t staged_tiles = aligned ? Q / 32 : 0;
  const int threads = 32 * kMpesWarps;
  constexpr int P = perm_count(C, K);
  const int smem = kMpesWarps * mpes_warp_smem_bytes(C, P);
  cudaFuncSetAttribute(mpes_all_perms_kernel<C, K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mpes_all_perms_kernel<C, K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (staged_tiles > 0) {
    const int64_t units = staged_tiles * pt.J;
    const int grid = static_cast<int>(std::min<int64_t>(4 * sm_count, (units + kMpesWarps - 1) / kMpesWarps));
    mpes_all_perms_kernel<C, K, true><<<grid, threads, smem, stream>>>(fchi, order, gs, pt.chunk_group, pt.J, S, Q, pt.Qpad, M,
                                                                   mode, 0, static_cast<int>(staged_tiles), pt.counters,
                                                                   pt.tile_done, pt.totbuf, pt.t1buf, out);
  }
  if (staged_tiles < tiles) {
    const int64_t units = (tiles - staged_tiles) * pt.J;
    const int grid = static_cast<int>(std::min<int64_t>(4 * sm_count, (units + kMpesWarps - 1) / kMpesWarps));
    mpes_all_perms_kernel<C, K, false><<<grid, threads, smem, stream>>>(fchi, order, gs, pt.chunk_group, pt.J, S, Q, pt.Qpad, M,
                                                                    mode, static_cast<int>(staged_tiles),
                                                                    static_cast<int>(tiles), pt.counters + 1, pt.tile_done,
                                                                    pt.totbuf, pt.t1buf, out);
  }
}

static bool dispatch_all_perms(int C, int K, const float* fchi, const int* order, const int* gs, int S, int64_t Q, int M,
                               int mode, int sm, const MpesParts& pt, double* out, cudaStream_t st) {
#define TKBO_CASE(c, k) \
  if (C == c && K == k) { launch_all_perms<c, k>(fchi, order, gs, S, Q, M, mode, sm, pt, out, st); return true; }
  TKBO_CASE(2, 1) TKBO_CASE(2, 2)
  TKBO_CASE(3, 1) TKBO_CASE(3, 2) TKBO_CASE(3, 3)
  TKBO_CASE(4, 1) TKBO_CASE(4, 2) TKBO_CASE(4, 3) TKBO_CASE(4, 4)
  TKBO_CASE(5, 1) TKBO_CASE(5, 2) TKBO_CASE(5, 3)
  TKBO_CASE(6, 1) TKBO_CASE(6, 2)
#undef TKBO_CASE
  return false;
}

static bool has_all_perms_kernel(int C, int K) {
  return (C == 2 && K <= 2) || (C == 3 && K <= 3) || (C == 4 && K <= 4) || (C == 5 && K <= 3) || (C == 6 && K <= 2);
}

static void full_table(int C, int k, std::vector<int>* out) {
  // lexicographic k-permutations, the order of rank_pes.get_all_permutation_k_in_n (rank_pes.py:218-228)
  std::vector<int> cur;
  std::vector<char> used(C, 0);
  struct Rec {
    static void go(int C, int k, std::vector<int>& cur, std::vector<char>& used, std::vector<int>* out) {
      if (static_cast<int>(cur.size()) == k) { out->insert(out->end(), cur.begin(), cur.end()); return; }
      for (int c = 0; c < C; ++c) {
        if (used[c]) continue;
        used[c] = 1; cur.push_back(c);
        go(C, k, cur, used, out);
        cur.pop_back(); used[c] = 0;
      }
    }
  };
  Rec::go(C, k, cur, used, out);
}

// Scratch of one call, allocated from the stream-ordered pool (cudaMallocAsync) so that calls on different streams or
// threads never share buffers; freed on the same stream when the object goes out of scope.
struct StreamScratch {
  void* p = nullptr;
  cudaStream_t stream;
  explicit StreamScratch(cudaStream_t s) : stream(s) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes, stream); }
  ~StreamScratch() { if (p) cudaFreeAsync(p, stream); }
};
static size_t align256(size_t b) { return (b + 255) / 256 * 256; }

int mpes_topk(tkbo_handle_t h, const float* fchi, const int32_t* fstar_argmax, int S, int64_t Q, int C, int M,
              const int32_t* perms, int P, int k, int mode, double* out_mi, cudaStream_t stream) {
  TKBO_REQUIRE(h && fchi && fstar_argmax && out_mi, "null pointer");
  TKBO_REQUIRE(S >= 1 && Q >= 1 && Q < (int64_t(1) << 31) && M >= 1, "S, Q, M must be >= 1 and Q < 2^31");
  TKBO_REQUIRE(C >= 1 && C <= kGenMaxC && k >= 1 && k <= C, "need 1 <= k <= C <= 8");
  TKBO_REQUIRE(mode == TKBO_MPES_RANK || mode == TKBO_MPES_PES, "unknown mode");
  if (mode == TKBO_MPES_PES) TKBO_REQUIRE(k == 1 && perms == nullptr, "PES mode is top-1 with the implicit table");
  const bool fast = perms == nullptr && has_all_perms_kernel(C, k);
  // layout of the call's scratch: order[S] | group_start[M + 1] | table[kGenMaxP * kGenMaxC] | chunk_group[kMaxChunks + 2] |
  // counters[2] | tile_done[tiles] | t1buf[J, Qpad] f64 | totbuf[J, P, Qpad] f32
  const int64_t tiles = (Q + 31) / 32;
  MpesParts pt;
  pt.Qpad = tiles * 32;
  pt.J = 1;
  const int Pn = fast ? perm_count(C, k) : 0;
  size_t per_chunk = 0;
  if (fast) {
    // sample chunks: enough (tile, chunk) units for ~8 rounds of the persistent warps, at most kMaxChunks, one per
    // maximiser, and partial buffers below 256 MiB; a large batch needs none (J = 1: no partial buffers at all)
    per_chunk = static_cast<size_t>(pt.Qpad) * (static_cast<size_t>(Pn) * sizeof(float) + sizeof(double));
    const int64_t warps = static_cast<int64_t>(h->sm_count) * 4 * kMpesWarps;
    int64_t J = (8 * warps + tiles - 1) / tiles;
    J = std::min<int64_t>(J, std::min(kMaxChunks, M));
    J = std::min<int64_t>(J, static_cast<int64_t>((size_t(256) << 20) / per_chunk));
    pt.J = static_cast<int>(std::max<int64_t>(1, J));
    if (const char* env = getenv("TKBO_MPES_CHUNKS")) {
      const int t = atoi(env);
      if (t >= 1 && t <= std::min(kMaxChunks, M)) pt.J = t;
    }
  }
  const size_t o_order = 0;
  const size_t o_gs = o_order + align256(sizeof(int) * S);
  const size_t o_table = o_gs + align256(sizeof(int) * (M + 1));
  const size_t o_cg = o_table + align256(fast ? 0 : sizeof(int) * kGenMaxP * kGenMaxC);
  const size_t o_cnt = o_cg + align256(sizeof(int) * (kMaxChunks + 2));
  const size_t o_done = o_cnt + 256;
  const size_t o_t1 = o_done + align256((fast && pt.J > 1) ? sizeof(unsigned int) * tiles : 0);
  const size_t o_tot = o_t1 + align256(pt.J > 1 ? sizeof(double) * pt.J * pt.Qpad : 0);
  const size_t total = o_tot + align256(pt.J > 1 ? sizeof(float) * pt.J * Pn * pt.Qpad : 0);
  StreamScratch sc(stream);
  TKBO_CUDA_CHECK(sc.alloc(total));
  unsigned char* base = static_cast<unsigned char*>(sc.p);
  int* order = reinterpret_cast<int*>(base + o_order);
  int* gs = reinterpret_cast<int*>(base + o_gs);
  int* dperm = reinterpret_cast<int*>(base + o_table);
  bool done = false;
  if (fast) {
    int* chunk_group = reinterpret_cast<int*>(base + o_cg);
    pt.chunk_group = chunk_group;
    pt.counters = reinterpret_cast<unsigned int*>(base + o_cnt);
    pt.tile_done = reinterpret_cast<unsigned int*>(base + o_done);
    pt.t1buf = reinterpret_cast<double*>(base + o_t1);
    pt.totbuf = reinterpret_cast<float*>(base + o_tot);
    TKBO_CUDA_CHECK(cudaMemsetAsync(base + o_cnt, 0, o_t1 - o_cnt, stream));      // unit counters + tile_done
    group_samples_kernel<<<1, 256, sizeof(int) * (M + 1), stream>>>(fstar_argmax, S, M, order, gs, pt.J, chunk_group);
    h->launches += 1;
    done = dispatch_all_perms(C, k, fchi, order, gs, S, Q, M, mode, h->sm_count, pt, out_mi, stream);
  } else {
    group_samples_kernel<<<1, 256, sizeof(int) * (M + 1), stream>>>(fstar_argmax, S, M, order, gs);
    h->launches += 1;
  }
  if (!done) {
    const int32_t* table = perms;
    int Pt = P;
    if (perms == nullptr) {
      std::vector<int> host;
      full_table(C, k, &host);
      Pt = static_cast<int>(host.size() / k);
      if (Pt > kGenMaxP) {
        set_error("all-permutation table has more than 1024 rows; pass an explicit (sub)table as rank_pes.py:49-53 does");
        return TKBO_ERR_UNSUPPORTED;
      }
      TKBO_CUDA_CHECK(cudaMemcpyAsync(dperm, host.data(), sizeof(int) * host.size(), cudaMemcpyHostToDevice, stream));
      TKBO_CUDA_CHECK(cudaStreamSynchronize(stream));     // host vector goes out of scope
      table = dperm;
    }
    TKBO_REQUIRE(Pt >= 1 && Pt <= kGenMaxP, "explicit table must have 1..1024 rows");
    const int64_t threads = Q * 32;
    mpes_table_kernel<<<static_cast<unsigned>((threads + 127) / 128), 128, Pt * kGenMaxC, stream>>>(
        fchi, order, gs, S, Q, C, M, table, Pt, k, mode, out_mi);
  }
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int mpes_indiff(tkbo_handle_t h, const float* fchi, const int32_t* fstar_argmax, int S, int64_t Q, int C, int M,
                double threshold, double* out_mi, cudaStream_t stream) {
  TKBO_REQUIRE(h && fchi && fstar_argmax && out_mi, "null pointer");
  TKBO_REQUIRE(S >= 1 && Q >= 1 && Q < (int64_t(1) << 31) && M >= 1, "S, Q, M must be >= 1 and Q < 2^31");
  TKBO_REQUIRE(C >= 1 && C <= kGenMaxC, "need 1 <= C <= 8");
  TKBO_REQUIRE(threshold >= 0.0 && threshold < 80.0, "indifference threshold must be in [0, 80)");
  StreamScratch sc(stream);
  TKBO_CUDA_CHECK(sc.alloc(align256(sizeof(int) * S) + sizeof(int) * (M + 1)));
  int* order = static_cast<int*>(sc.p);
  int* gs = reinterpret_cast<int*>(static_cast<unsigned char*>(sc.p) + align256(sizeof(int) * S));
  group_samples_kernel<<<1, 256, sizeof(int) * (M + 1), stream>>>(fstar_argmax, S, M, order, gs);
  mpes_indiff_kernel<<<static_cast<unsigned>((Q + 127) / 128), 128, 0, stream>>>(fchi, order, gs, S, Q, C, M,
                                                                            static_cast<float>(exp(threshold)), out_mi);
  h->launches += 2;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int argmax_rows(tkbo_handle_t h, const float* fstar, int S, int M, int32_t* out, cudaStream_t stream) {
  TKBO_REQUIRE(h && fstar && out && S >= 1 && M >= 1, "bad argmax_rows arguments");
  argmax_rows_f32_kernel<<<(S + 127) / 128, 128, 0, stream>>>(fstar, S, M, out);
  h->launches += 1;
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

int soft_copeland(tkbo_handle_t h, const float* f, int n, float* out_score, int64_t* out_argmax, cudaStream_t stream) {
  TKBO_REQUIRE(h && f && n >= 1 && (out_score || out_argmax), "bad soft_copeland arguments");
  float* score = out_score;
  if (!score) TKBO_CUDA_CHECK(cudaMallocAsync(&score, sizeof(float) * n, stream));
  soft_copeland_rows_kernel<<<n, 256, 0, stream>>>(f, n, score);
  h->launches += 1;
  if (out_argmax) {
    argmax_vec_kernel<<<1, 256, 0, stream>>>(score, n, reinterpret_cast<long long*>(out_argmax));
    h->launches += 1;
  }
  if (!out_score) TKBO_CUDA_CHECK(cudaFreeAsync(score, stream));
  TKBO_CUDA_CHECK(cudaGetLastError());
  return TKBO_OK;
}

}  // namespace tkbo
