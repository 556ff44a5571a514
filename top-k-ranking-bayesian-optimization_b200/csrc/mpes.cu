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
// group_const[2 m] = p_m (rank_pes.py:40-41; pes.py:66-70 with the eps smoothing), group_const[2 m + 1] = log(1 / (S p_m)).
// `zero` (optional): 64 words cleared for the kernel that follows (its tile counters), which saves a memset node per call.
__global__ void group_samples_kernel(const int* __restrict__ arg, int S, int M, int* __restrict__ order,
                                     int* __restrict__ group_start, int mode = 0, double* __restrict__ group_const = nullptr,
                                     unsigned int* __restrict__ zero = nullptr) {
  extern __shared__ int cnt[];           // [M + 1]
  if (zero != nullptr && threadIdx.x < 64) zero[threadIdx.x] = 0u;
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
  if (group_const != nullptr) {
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
      const double c = static_cast<double>(cnt[m + 1] - cnt[m]);
      const double pm = (mode == TKBO_MPES_PES) ? (c + kPesEps) / (static_cast<double>(S) + M * kPesEps) : c / static_cast<double>(S);
      group_const[2 * m] = pm;
      group_const[2 * m + 1] = (pm > 0.0) ? log(1.0 / (static_cast<double>(S) * pm)) : 0.0;
    }
  }
  // stable scatter by one warp, 32 samples at a time: a sample's slot = its group's cursor + the number of lower lanes
  // with the same winner (match.any); the lowest lane of each set advances the cursor.  (One thread per group scanning all
  // S samples took ~20 us at S = 1024, 2 % of the c5 call.)
  __syncthreads();                        // cnt[m] has been read as p_m above; from here on it is the cursor of group m
  if (threadIdx.x < 32) {
    const unsigned lane = threadIdx.x;
    for (int s0 = 0; s0 < S; s0 += 32) {
      const int sidx = s0 + static_cast<int>(lane);
      const int m = (sidx < S) ? arg[sidx] : -1;
      const bool ok = (m >= 0 && m < M);
      const unsigned peers = __match_any_sync(0xffffffffu, ok ? m : -1 - static_cast<int>(lane));
      if (ok) {
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int basepos = cnt[m];
        order[basepos + rank] = sidx;
        __syncwarp(peers);                // every peer has read the cursor
        if (rank == 0) cnt[m] = basepos + __popc(peers);
      }
      __syncwarp();
    }
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
// The sum over a set is the sum over the set without its lowest member plus that member, so every set costs one addition
// (26 for C = 5; the compiler drops the sets nothing needs).
__host__ __device__ constexpr int lowest_bit_c(unsigned m) {
  int c = 0;
  for (; !((m >> c) & 1u); ++c) {}
  return c;
}
template <int C, int K>
__device__ __forceinline__ void subset_reciprocals(const float (&e)[C], float (&rs)[1 << C]) {
  float ss[1 << C];                                        // ss[set] = sum of e over the set
#pragma unroll
  for (unsigned set = 1; set < (1u << C); ++set) {
    const unsigned rest = set & (set - 1u);
    ss[set] = (rest == 0u) ? e[lowest_bit_c(set)] : ss[rest] + e[lowest_bit_c(set)];
  }
#pragma unroll
  for (unsigned m = 0; m < (1u << C); ++m) {
    if (popcount_c(m) >= K || popcount_c(m) >= C - 1) continue;       // a lone remaining choice has probability 1
    rs[m] = fast_rcp(ss[((1u << C) - 1u) & ~m]);          // > 0: the callers keep every e[c] >= 2^-126
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

// Packed walk (K >= 2): sm_100's two-wide fp32 instructions (FMUL2 / FFMA2, one issue slot for two results) with the two
// halves working on mirror images of each other.  Relabelling the choices c -> C-1-c maps orderings onto orderings without
// fixed points, so the walk visits only the prefixes that are lexicographically <= their mirror image, carries
// (prob(prefix), prob(mirror prefix)) and multiplies by (e_c, e_{C-1-c}) and (rs[used], rs[mirror(used)]): half the
// multiply / accumulate instructions of PermWalk for the same P results.  acc[j] = (ordering j, its mirror image).
__host__ __device__ constexpr unsigned mask_mirror(unsigned m, int C) {
  unsigned r = 0;
  for (int c = 0; c < C; ++c)
    if ((m >> c) & 1u) r |= 1u << (C - 1 - c);
  return r;
}
template <int C, int K, int DEPTH>
struct PermWalk2 {
  // tied: the prefix so far equals its own mirror image (empty, or the middle choice of an odd C)
  template <int P2>
  static __device__ __forceinline__ void run(const float2 (&e2)[C], const float (&rs)[1 << C], unsigned used, bool tied,
                                             float2 prob, float2 (&acc)[P2], int& leaf) {
    constexpr int remaining = C - DEPTH;
    float2 pr = prob;
    if constexpr (remaining > 1) {
      const float2 r2 = make_float2(rs[used], rs[mask_mirror(used, C)]);
      pr = (DEPTH == 0) ? r2 : __fmul2_rn(prob, r2);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if ((used >> c) & 1u) continue;
      if (tied && c > C - 1 - c) continue;                 // its mirror image is the one that is walked
      if constexpr (DEPTH + 1 == K) {
        acc[leaf] = (remaining > 1) ? __ffma2_rn(pr, e2[c], acc[leaf]) : __fadd2_rn(acc[leaf], pr);
        ++leaf;
      } else {
        const float2 child = (remaining > 1) ? __fmul2_rn(pr, e2[c]) : pr;
        PermWalk2<C, K, DEPTH + 1>::run(e2, rs, used | (1u << c), tied && (c == C - 1 - c), child, acc, leaf);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// fast kernel: one thread per query, all K-permutations of C in registers.  A block of kMpesWarps warps owns one tile of
// 32 queries at a time (tiles are handed out by a global counter); its warps split the SAMPLES of the tile: warp w takes
// positions w, w + W, w + 2W, ... of the grouped sample order, whatever group they fall in, so the split is even however
// skewed the groups are (a maximiser that wins 70 % of the samples is the normal case late in a BO run).  Each warp streams
// its samples' rows (32*C contiguous floats per sample) through a private shared-memory ring filled kRing rows ahead by
// cp.async copies: 16 bytes per lane and instruction (two coalesced LDGSTS per row and warp; the 1-D TMA bulk copies of
// round 1 cost ~40 instructions of elect / mbarrier / descriptor work per row, a sixth of the loop), or 4 bytes (C
// instructions per row) when the rows are not multiples of 16 bytes (template CPB).  A ragged last tile copies only the
// part of its row segment that exists.
// At every group boundary the warps add their partial sums into a block-shared accumulator and the group is folded once.
// ------------------------------------------------------------------------------------------------
// Rows in flight per warp (power of two).  Few choices = short rows (C = 2: 256 bytes per row and warp): with 4 rows ahead
// 16 warps hold 16 KB in flight per SM and the pairwise shape - the reference's default - stopped at 2.3 TB/s, bound by
// bytes in flight, not by arithmetic; ~4 KB per warp are needed.  C = 5 is bound by its arithmetic and keeps 4.
__host__ __device__ constexpr int ring_rows(int C) { return C <= 2 ? 16 : (C <= 4 ? 8 : 4); }
#ifndef TKBO_MPES_WARPS
#define TKBO_MPES_WARPS 4           // measured on B200 at c5 (skewed groups, same box): 4 warps 1.25 ms, 8 warps 1.34 ms
#endif
constexpr int kMpesWarps = TKBO_MPES_WARPS;
// blocks per SM: 16 warps of <= 128 registers fill the register file.  Five blocks of four warps (96 registers, 32 bytes of
// spills) measured slower at c5: 1.072 against 1.040 ms
#ifndef TKBO_MPES_BLOCKS
#define TKBO_MPES_BLOCKS (16 / TKBO_MPES_WARPS)
#endif
constexpr int kMpesBlocksPerSm = TKBO_MPES_BLOCKS;    // 16 warps of 128 registers fill the register file
__host__ __device__ constexpr int mpes_warp_smem_bytes(int C) {
  return ((ring_rows(C) * 32 * C * 4 + 127) / 128) * 128;
}
// one row of 32 * C floats (16-byte aligned) -> shared memory, 16 bytes per lane and instruction; one commit group per row.
// dst_lane / src_lane already point at this lane's first 16-byte chunk; a row has 8 C <= 64 chunks.
// One row segment of a tile (32 C floats, fewer in a ragged last tile) -> shared memory; one commit group per row.
// CPB = 16: rows that are multiples of 16 bytes from a 16-byte aligned base, one or two coalesced 16-byte copies per lane;
// CPB = 4: any other row length, C coalesced 4-byte copies per lane.  `units` = CPB-sized pieces of the segment that lie
// inside the row, so nothing is read past the row or written past the ring slot; dst_lane / src_lane point at this lane's
// first piece.
template <int C, int CPB>
__device__ __forceinline__ void row_copy_async(uint32_t dst_lane, const float* src_lane, int lane, int units) {
  if constexpr (CPB == 16) {
    if (lane < units)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_lane), "l"(src_lane) : "memory");
    if (lane + 32 < units)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_lane + 512u), "l"(src_lane + 128) : "memory");
  } else {
#pragma unroll
    for (int t = 0; t < C; ++t)
      if (lane + 32 * t < units)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_lane + 128u * t), "l"(src_lane + 32 * t) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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
  double rd[128];     // r[k] 2^-31: for integers normalised to [2^31, 2^32)
};
__device__ __forceinline__ void log_table_fill(LogTable& tab, int tid) {
  if (tid < 128) {
    const float r0 = 1.f / (1.f + (static_cast<float>(tid) + 0.5f) * (1.f / 128.f));
    const float r = __uint_as_float(__float_as_uint(r0) & 0xFFFFC000u);       // 9 explicit mantissa bits
    tab.r[tid] = r;
    tab.t[tid] = -log(static_cast<double>(r));
    tab.rd[tid] = static_cast<double>(r) * (1.0 / 2147483648.0);
  }
}
// log(n) for an integer n >= 1, all in float64 (about 1e-15): n = 2^E v 2^-31 with v in [2^31, 2^32), k = the 7 bits below
// v's leading one, x = v rd[k] - 1 exactly (32 x 9 bits), log1p(x) = x (1 - x/2 + x^2/3 - x^3/4) up to x^5/5 < 2e-12.
// exact uint32 -> float64 on the float64 pipe (2^52 + n, minus 2^52): the I2F.F64 conversions run on the 16-lane XU
// pipe, which the sample loop already saturates
__device__ __forceinline__ double u32_to_double(unsigned int n) {
  return __hiloint2double(0x43300000, static_cast<int>(n)) - 4503599627370496.0;
}
__device__ __forceinline__ double table_log_uint(const LogTable& tab, unsigned int n) {
  const int lz = __clz(static_cast<int>(n));
  const unsigned int v = n << lz;
  const unsigned int k = (v >> 24) & 127u;
  const double x = fma(u32_to_double(v), tab.rd[k], -1.0);
  const double p = fma(x, fma(x, fma(x, -0.25, 0.3333333333333333), -0.5), 1.0);
  return fma(u32_to_double(static_cast<unsigned int>(31 - lz)), 0.6931471805599453, fma(x, p, tab.t[k]));
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

// Per group m the block needs  a_z = sum_{s in m} P_s(z)  over ALL its warps' samples.  Every warp rounds its partial sum
// to a 2^-F grid (F = 31 - ceil(log2 S): sums are at most S) and adds it, as an integer, into gacc[z][lane] with a shared
// memory reduction: integer sums commute, so the result is bitwise independent of warp timing.  After one block barrier
// warp w folds the orderings z = w, w + W, ...:
//     p_m sum_z c log c = (1/S) (sum_z a_z log a_z + log(cm) sum_z a_z),   c = a cm,  cm = 1 / (S p_m)
// (float64 sums, table_log above) and keeps  tot_z = sum_m a_z  for its orderings in registers - again as integers.
// Consistency of p(z) with the joint matters: MI = sum J log(J / (p_m p_z)) is insensitive to first order to errors in J
// only if p_z is EXACTLY sum_m J_zm; a float32 running sum of the group values breaks that at the 1e-7 level, which is
// 1e-4 of a typical MI.  The grid value is what the fold uses AND what is summed into tot_z.
// At the end of the tile  MI = sum_w A_w  with  A_w = sum_m (fold terms of w's orderings) - sum_{z of w} p_z log p_z.
template <int C, int K, int CPB /* bytes per staged copy: 16 or 4 */>
__global__ void __launch_bounds__(32 * kMpesWarps, kMpesBlocksPerSm)
mpes_all_perms_kernel(const float* __restrict__ fchi, const int* __restrict__ order, const int* __restrict__ group_start,
                      const double* __restrict__ group_const, int S, long long Q, int M, int mode, int tile_begin,
                      int tile_end, unsigned int* __restrict__ tile_counter, double* __restrict__ out_mi) {
  constexpr int P = perm_count(C, K);
  constexpr int W = kMpesWarps;
  constexpr int kOwn = (P + W - 1) / W;                    // orderings folded by one warp
  constexpr int kRing = ring_rows(C);
  constexpr bool kPacked = (K >= 2);                       // PermWalk2 (P is even then)
  constexpr float kLog2e = 1.4426950408889634f;
  // exponent offset: e_c = 2^((f_c - max) log2e + 120), so that choices up to 246 binades (170 nats) below the best one stay
  // normal floats (the sums stay below 2^127, every stage factor e * rs is scale free); rank_pes.py:166 renormalises per
  // stage, which this reproduces as long as the not-yet-ranked choices do not ALL flush to zero
  constexpr float kExpBias = 120.f;
  __shared__ LogTable logtab;
  __shared__ unsigned int gacc[2][P][32];                  // block-wide group sums (fixed point), double-buffered by group
  __shared__ double a_red[W][32];
  __shared__ uint64_t gbar[2];                             // "every warp has delivered its share" per gacc buffer
  __shared__ unsigned int tot_sh[(P + W - 1) / W][W][32];  // tot_z = sum_m a_z of the orderings a warp owns (fixed point)
  __shared__ int tile_sh;
  // dynamic shared memory, one private slice per warp: the row ring (see mpes_warp_smem_bytes)
  extern __shared__ __align__(128) unsigned char mpes_smem[];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  log_table_fill(logtab, threadIdx.x);
  for (int i = threadIdx.x; i < 2 * P * 32; i += blockDim.x) (&gacc[0][0][0])[i] = 0u;
  unsigned char* slice = mpes_smem + static_cast<size_t>(wib) * mpes_warp_smem_bytes(C);
  float (*ring)[32 * C] = reinterpret_cast<float (*)[32 * C]>(slice);
  const float* ring_lane = &ring[0][lane * C];             // this lane's C values of slot 0; slot s is 32 C floats further
  const uint32_t ring_dst = smem_u32(&ring[0][(CPB / 4) * lane]);  // this lane's first piece of slot 0
  const unsigned int rs32 = static_cast<unsigned int>(Q * C);          // floats per sample row (host checks Q C < 2^31)
  const double invS = 1.0 / static_cast<double>(S);
  const double eps = (mode == TKBO_MPES_PES) ? kPesEps : 0.0;
  const double meps = (mode == TKBO_MPES_PES) ? M * kPesEps : 0.0;
  int fix_bits = 31;
  while (fix_bits > 0 && (static_cast<long long>(S) << fix_bits) > (1ll << 31)) --fix_bits;   // S 2^F <= 2^31
  const float fix_scale = __int_as_float((127 + fix_bits) << 23);                              // 2^F
  const double fix_unit = 1.0 / static_cast<double>(1ll << fix_bits);                          // 2^-F
  const int i1 = group_start[M];                           // samples with a valid winner
  if (threadIdx.x == 0) {
    mbar_init(&gbar[0], W);
    mbar_init(&gbar[1], W);
  }
  fence_mbar_init();
  uint32_t g = 0;                                          // groups delivered so far by this block (runs across tiles)

  for (;;) {
    if (threadIdx.x == 0) tile_sh = tile_begin + static_cast<int>(atomicAdd(tile_counter, 1u));
    __syncthreads();                                       // tile_sh visible; a_red / gacc of the previous tile are done with
    const int tile = tile_sh;
    if (tile >= tile_end) break;
    const long long q = static_cast<long long>(tile) * 32 + lane;
    const bool valid = q < Q;
    const float* tile_base = fchi + static_cast<long long>(tile) * 32 * C + (CPB / 4) * lane;   // this lane's first piece
    // pieces of a row segment that lie inside the row: all of them except in a ragged last tile
    const int chunks = static_cast<int>(min(static_cast<long long>(32 * C * 4 / CPB),
                                            (Q * C - static_cast<long long>(tile) * 32 * C) / (CPB / 4)));
    float acc[kPacked ? 1 : P];                            // K = 1: one accumulator per ordering
    float2 acc2[kPacked ? P / 2 : 1];                      // K >= 2: (ordering j, its mirror image = ordering j + P/2)
#pragma unroll
    for (int z = 0; z < (kPacked ? 1 : P); ++z) acc[z] = 0.f;
#pragma unroll
    for (int z = 0; z < (kPacked ? P / 2 : 1); ++z) acc2[z] = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 0; t < kOwn; ++t) tot_sh[t][wib][lane] = 0u;          // (private to this thread: no synchronisation needed)
    double A = 0.0;

    {                                                      // prime the ring with this warp's first kRing rows
      __syncwarp();                                        // (the previous tile's last rows have been consumed by every lane)
#pragma unroll
      for (int r = 0; r < kRing; ++r) {                    // one commit group per slot, empty past the end
        if (wib + r * W < i1)
          row_copy_async<C, CPB>(ring_dst + r * (32 * C * 4), tile_base + static_cast<size_t>(static_cast<unsigned int>(order[wib + r * W])) * rs32, lane, chunks);
        cp_async_commit();
      }
    }
    // sample index of the row the NEXT refill fetches, loaded one iteration early: the refill's address would otherwise wait
    // for this load in every iteration (the top stall of the short-row shapes)
    unsigned int ord_ahead = static_cast<unsigned int>(order[min(wib + kRing * W, i1 > 0 ? i1 - 1 : 0)]);

    int i = wib;                                           // position in the grouped sample order
    unsigned int n = 0;                                    // samples this warp has consumed (ring slot = n % kRing)
    // Software pipeline over the groups: a warp delivers its share of group m and goes straight on to the samples of the
    // next group; it folds group m only after those (by then the other warps have delivered too, so the wait on gbar is
    // normally already satisfied).  Order per warp: samples(m), [wait + fold(previous)], deliver(m), arrive.  Buffer reuse:
    // deliver(m + 2) follows this warp's wait for group m + 1, i.e. every warp's arrive(m + 1), which each warp issues after
    // its fold(m) has read and cleared the buffer.
    int prev = -1;                                         // group delivered but not folded yet
    for (int m = 0; m <= M; ++m) {
      const bool tail = (m == M);                          // one extra pass to fold the last group
      const unsigned int n_before = n;
      if (!tail) {
        const int hi = group_start[m + 1];
        if (hi == group_start[m] && mode == TKBO_MPES_RANK) continue;   // rank_pes.py:86-93 drops empty groups (block-uniform)
        // (Tried in round 2: the loads / max / exponentials of sample n + 1 issued inside the iteration of sample n - a
        // software pipeline by hand.  128 registers with spills, 1.112 against 1.034 ms at c5: the four warps of a
        // sub-partition already interleave these phases.)
        for (; i < hi; i += W, ++n) {
          float f[C];
          {
            const int slot = n % kRing;
            cp_async_wait<kRing - 1>();                    // this lane's copies of row n have landed ...
            __syncwarp();                                  // ... and so have the other lanes'
#pragma unroll
            for (int c = 0; c < C; ++c) f[c] = ring_lane[slot * (32 * C) + c];
          }
          float mx = f[0];
#pragma unroll
          for (int c = 1; c < C; ++c) mx = fmaxf(mx, f[c]);
          const float mxs = fmaf(mx, kLog2e, -kExpBias);
          float e[C];
#pragma unroll
          // floor at the smallest normal: a set of remaining choices that has underflowed entirely (> 170 nats below the
          // best) then shares its (negligible) mass evenly instead of producing 0 * inf
          for (int c = 0; c < C; ++c) e[c] = fmaxf(fast_ex2(fmaf(f[c], kLog2e, -mxs)), 1.17549435e-38f);
          {
            __syncwarp();                                  // every lane has consumed its slot values
            if (i + kRing * W < i1)
              row_copy_async<C, CPB>(ring_dst + (n % kRing) * (32 * C * 4), tile_base + static_cast<size_t>(ord_ahead) * rs32, lane, chunks);
            cp_async_commit();                             // (an empty group past the end keeps the group count in step)
            ord_ahead = static_cast<unsigned int>(order[min(i + (kRing + 1) * W, i1 - 1)]);
          }
          float rs[1 << C];
          subset_reciprocals<C, K>(e, rs);
          int leaf = 0;
          if constexpr (kPacked) {
            float2 e2[C];
#pragma unroll
            for (int c = 0; c < C; ++c) e2[c] = make_float2(e[c], e[C - 1 - c]);
            PermWalk2<C, K, 0>::run(e2, rs, 0u, true, make_float2(1.f, 1.f), acc2, leaf);
          } else {
            PermWalk<C, K, 0>::run(e, rs, 0u, 1.f, acc, leaf);
          }
        }
      }
      if (prev >= 0) {
        // fold the orderings this warp owns of group `prev`:  A += p_m sum_z c_zm log c_zm,  c_zm = p(z | m) = (a_z + eps) cm
        const uint32_t gp = g - 1u;
        const int buf = static_cast<int>(gp & 1u);
        mbar_wait(&gbar[buf], (gp >> 1) & 1u);
        const double lcm = group_const[2 * prev + 1];      // log cm = log(1 / (S p_m))
        double s1 = 0.0, s0 = 0.0;
        if (mode == TKBO_MPES_RANK) {
          // eps = 0: a_z = fx 2^-F, so  sum a log a = 2^-F (sum fx log fx - F ln2 sum fx)  with an all-float64 integer log
          unsigned int isum = 0u;
#pragma unroll
          for (int t = 0; t < kOwn; ++t) {
            const int z = wib + t * W;
            if (z < P) {
              const unsigned int fx = gacc[buf][z][lane];
              gacc[buf][z][lane] = 0u;                     // free for the group after next
              // branch-free (0 log 1 = 0): the P / W logarithms of a fold are independent chains the compiler can interleave
              s1 = fma(u32_to_double(fx), table_log_uint(logtab, fx | static_cast<unsigned int>(fx == 0u)), s1);
              isum += fx;                                  // <= S 2^F <= 2^31 over ALL orderings
              tot_sh[t][wib][lane] += fx;
            }
          }
          s0 = static_cast<double>(isum) * fix_unit;
          s1 = fma(-static_cast<double>(fix_bits) * 0.6931471805599453, static_cast<double>(isum), s1) * fix_unit;
        } else {
#pragma unroll
          for (int t = 0; t < kOwn; ++t) {
            const int z = wib + t * W;
            if (z < P) {
              const unsigned int fx = gacc[buf][z][lane];
              gacc[buf][z][lane] = 0u;
              // a = grid sum + eps in float64 (eps = 1e-7 is below float32 resolution of a sum of ~S/M probabilities, and
              // pes.I of an uninformative query consists of nothing but these eps terms):
              // log a = log(fl32(a)) + (a - fl32(a)) / fl32(a)
              const double ad = u32_to_double(fx) * fix_unit + eps;
              const float af0 = static_cast<float>(ad);
              const bool tiny = !(af0 >= 1e-30f);          // a log a < 7e-29: dropped (and no denormal handling needed)
              const float af = tiny ? 1.f : af0;           // branch-free: independent chains
              const double adk = tiny ? 0.0 : ad;
              const double la = table_log(logtab, af) + (adk - static_cast<double>(af)) * static_cast<double>(fast_rcp(af));
              s1 = fma(adk, tiny ? 0.0 : la, s1);
              s0 += adk;
              tot_sh[t][wib][lane] += fx;
            }
          }
        }
        A += invS * (s1 + lcm * s0);
        prev = -1;
      }
      if (tail) break;
      // deliver this warp's share of group m (integers: the block sum does not depend on the order of arrival); a warp
      // that saw no sample of the group (groups smaller than the block) only arrives
      const int buf = static_cast<int>(g & 1u);
      if (n != n_before) {
        if constexpr (kPacked) {
#pragma unroll
          for (int z = 0; z < P / 2; ++z) {
            atomicAdd(&gacc[buf][z][lane], __float2uint_rn(acc2[z].x * fix_scale));
            atomicAdd(&gacc[buf][z + P / 2][lane], __float2uint_rn(acc2[z].y * fix_scale));
            acc2[z] = make_float2(0.f, 0.f);
          }
        } else {
#pragma unroll
          for (int z = 0; z < P; ++z) {
            atomicAdd(&gacc[buf][z][lane], __float2uint_rn(acc[z] * fix_scale));
            acc[z] = 0.f;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&gbar[buf]);
      prev = m;
      ++g;
    }
    // - sum_z p_z log p_z over this warp's orderings, then the block sum of the W partial results
    if (mode == TKBO_MPES_RANK) {      // p_z = tot 2^-F / S: integer log again
      const double shift = static_cast<double>(fix_bits) * 0.6931471805599453 + log(static_cast<double>(S));
      double acc_t = 0.0;
      unsigned int isum = 0u;
#pragma unroll
      for (int t = 0; t < kOwn; ++t) {
        const unsigned int tz = (wib + t * W < P) ? tot_sh[t][wib][lane] : 0u;
        acc_t = fma(u32_to_double(tz), table_log_uint(logtab, tz | static_cast<unsigned int>(tz == 0u)), acc_t);
        isum += tz;
      }
      A -= (acc_t - shift * static_cast<double>(isum)) * fix_unit * invS;
    } else {
#pragma unroll
      for (int t = 0; t < kOwn; ++t) {
        if (wib + t * W < P) {
          const double pz = (static_cast<double>(tot_sh[t][wib][lane]) * fix_unit + meps) * invS;
          if (pz > 0.0) A -= pz * log(pz);
        }
      }
    }
    a_red[wib][lane] = A;
    __syncthreads();
    if (wib == 0 && valid) {
      double mi = 0.0;
#pragma unroll
      for (int w = 0; w < W; ++w) mi += a_red[w][lane];
      out_mi[q] = mi;
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
  float acc[32];
  double tot[32];                                   // exact sum of the group values: p_z must equal sum_m J_zm (see above)
#pragma unroll
  for (int i = 0; i < 32; ++i) { acc[i] = 0.f; tot[i] = 0.0; }
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
        tot[t] += static_cast<double>(acc[t]);
      }
      acc[t] = 0.f;
    }
  }
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    if (lane + 32 * t < P) {
      const double pz = (tot[t] + M * eps) * invS;
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
  float acc[kGenMaxC + 1];
  double tot[kGenMaxC + 1];                         // exact sum of the group values (p_z == sum_m J_zm)
#pragma unroll
  for (int z = 0; z <= kGenMaxC; ++z) { acc[z] = 0.f; tot[z] = 0.0; }
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
        tot[z] += static_cast<double>(acc[z]);
      }
      acc[z] = 0.f;
    }
  }
#pragma unroll
  for (int z = 0; z <= kGenMaxC; ++z) {
    if (z < C || z == kGenMaxC) {
      const double pz = tot[z] * invS;
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
template <int C, int K>
static void launch_all_perms(const float* fchi, const int* order, const int* gs, const double* gconst, int S, int64_t Q, int M,
                             int mode, int sm_count, unsigned int* counters, double* out, cudaStream_t stream) {
  const int64_t tiles = (Q + 31) / 32;
  // 16-byte copies need rows that are multiples of 16 bytes from an aligned base; anything else takes 4-byte copies (same
  // kernel otherwise: at c5, Q = 100001 instead of 100000, the former direct-load path took 1.36 ms against 1.04)
  const bool aligned = ((Q * C) % 4 == 0) && ((reinterpret_cast<uintptr_t>(fchi) & 15) == 0);
  const int threads = 32 * kMpesWarps;
  const int smem = kMpesWarps * mpes_warp_smem_bytes(C);
  const int grid = static_cast<int>(std::min<int64_t>(kMpesBlocksPerSm * sm_count, tiles));
  if (aligned) {
    cudaFuncSetAttribute(mpes_all_perms_kernel<C, K, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    mpes_all_perms_kernel<C, K, 16><<<grid, threads, smem, stream>>>(fchi, order, gs, gconst, S, Q, M, mode, 0,
                                                                 static_cast<int>(tiles), counters, out);
  } else {
    cudaFuncSetAttribute(mpes_all_perms_kernel<C, K, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    mpes_all_perms_kernel<C, K, 4><<<grid, threads, smem, stream>>>(fchi, order, gs, gconst, S, Q, M, mode, 0,
                                                                static_cast<int>(tiles), counters, out);
  }
}

static bool dispatch_all_perms(int C, int K, const float* fchi, const int* order, const int* gs, const double* gconst, int S,
                               int64_t Q, int M, int mode, int sm, unsigned int* counters, double* out, cudaStream_t st) {
#define TKBO_CASE(c, k) \
  if (C == c && K == k) { launch_all_perms<c, k>(fchi, order, gs, gconst, S, Q, M, mode, sm, counters, out, st); return true; }
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
  TKBO_REQUIRE(Q * C < (int64_t(1) << 31), "Q * C must stay below 2^31");
  TKBO_REQUIRE(mode == TKBO_MPES_RANK || mode == TKBO_MPES_PES, "unknown mode");
  if (mode == TKBO_MPES_PES) TKBO_REQUIRE(k == 1 && perms == nullptr, "PES mode is top-1 with the implicit table");
  const bool fast = perms == nullptr && has_all_perms_kernel(C, k);
  // layout of the call's scratch: order[S] | group_start[M + 1] | group_const[2 M] f64 | table[kGenMaxP * kGenMaxC] | tile counters[2]
  const size_t o_order = 0;
  const size_t o_gs = o_order + align256(sizeof(int) * S);
  const size_t o_gc = o_gs + align256(sizeof(int) * (M + 1));
  const size_t o_table = o_gc + align256(sizeof(double) * 2 * M);
  const size_t o_cnt = o_table + align256(fast ? 0 : sizeof(int) * kGenMaxP * kGenMaxC);
  const size_t total = o_cnt + 256;
  StreamScratch sc(stream);
  TKBO_CUDA_CHECK(sc.alloc(total));
  unsigned char* base = static_cast<unsigned char*>(sc.p);
  int* order = reinterpret_cast<int*>(base + o_order);
  int* gs = reinterpret_cast<int*>(base + o_gs);
  int* dperm = reinterpret_cast<int*>(base + o_table);
  double* gconst = reinterpret_cast<double*>(base + o_gc);
  group_samples_kernel<<<1, 256, sizeof(int) * (M + 1), stream>>>(fstar_argmax, S, M, order, gs, mode, gconst,
                                                              reinterpret_cast<unsigned int*>(base + o_cnt));   // + tile counters = 0
  h->launches += 1;
  bool done = false;
  if (fast) {
    done = dispatch_all_perms(C, k, fchi, order, gs, gconst, S, Q, M, mode, h->sm_count,
                              reinterpret_cast<unsigned int*>(base + o_cnt), out_mi, stream);
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
