// K1 — fused random-Fourier-feature evaluation, tcgen05 contraction with the posterior weight samples and
// per-sample (value, index) argmax.  Replaces fourier_features.py:18-21 + ":162-167" (n_init -> N) and the
// Phi @ theta of acquisitions/dts.py:82-85 for a basis shared by all S samples.
//
// Work item = (128-candidate tile, chunk of SC samples).  Per item the CTA accumulates
//     acc[128 x SC] = sum over 64-feature slabs of  Phi_slab[128 x 64] * ThetaScaled_slab[SC x 64]^T
// with Phi and Theta each split into fp16 hi + lo parts and three tcgen05.mma (hi*hi + hi*lo + lo*hi) per
// K=16 step, fp32 accumulation in TMEM ("fp32 emulation": ~2^-22 relative per product) -- or, template C8, hi*hi in
// fp16 plus ONE e4m3 tcgen05.mma (K = 32) that carries both correction terms: two thirds of the tensor work at ~3e-5
// of max|f| (see cos_split16_c8).
//
// Both GEMMs of the path run on the tensor pipe.  The projection  arg = X W^T + b  of a slab is itself a small
// tcgen05.mma: X and (W | b) are split into three bf16 terms each (x = x1 + x2 + x3 exactly) and the six products
// of order >= 2^-16 are laid side by side along K (K = 6 d + 3 <= 112), so that arg lands in TMEM with fp32
// accuracy.  The feature producers then only do  TMEM -> cos -> fp16 hi/lo split -> TMEM  in place: four to six
// instructions per feature, independent of d, no shared-memory traffic.
//
// Warp roles (32 * (6 + 4 NPG) threads, one CTA per SM, persistent over work items):
//   warps 0-3   per item: write the bf16-split candidate rows (projection A operand) into TMEM ahead of time;
//               epilogue: accumulator -> one compare per value against the running per-sample maxima (thr_s); only
//               32-column blocks that may improve a sample run the warp max (CREDUX) + ballot row + atomicMax of the
//               packed (value, ~row) key; or store F; or soft-Copeland sums (duel mode)
//   warps 4..   feature producers (NPG groups of four warps; group g takes global slab iterations g, g+NPG, ...):
//               thread = one candidate row: tcgen05.ld 64 projection values, cos, hi/lo split, tcgen05.st back
//               into the same 64 TMEM columns, which the main MMAs then read as their A operand
//   next warp   TMA: Theta hi/lo slab tiles (128B-swizzled, K-major) + the bf16 (W | b) slab tile into the smem ring
//               (wstages > 0: the (W | b) tiles have their own, deeper ring and run ahead of the Theta stages)
//   2 warps     projection MMAs, alternating slabs (1 warp when the smem ring depth is odd): they run ahead of the
//               main MMAs as far as free TMEM ring slots and landed tiles allow
//   2 warps     main MMAs, two slabs per burst when the operands are ready.  Sample chunks > 128 (single CTAs): each warp
//               owns one half of the accumulator columns (N = SC / 2 per instruction), the two warps' bursts offset by
//               one slab; chunks <= 128 and CTA pairs: one warp issues N = SC and the other idles (RffParams::nsplit).
//   A lone warp needs ~300 cycles for the waits, commits and address arithmetic of one slab while the tensor
//   pipe queues only ~4 instructions (tools/probe_mma_chain.cu); splitting the issue work by columns keeps every
//   accumulator column's MMAs in one program order, but halving N does not halve an MMA's cost, so it pays only for
//   wide chunks.
//
// CTA pairs (template PAIR, clusters of two, cta_group::2 MMAs with M = 256 issued by the leader; a second set of four
// epilogue warps; the single accumulator retired in two column halves): see the comment above the kernel.
//
// Around the launch: the last CTA unpacks the running keys and, for the multi-GPU merge, stores them into every rank's
// peer-mapped buffer and releases a flag (RffParams::peer_buf); with host candidates the row-staging warps acquire a
// "chunks landed" counter that the H2D copy stream bumps, so one launch overlaps the transfer (RffParams::ready_flag).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace tkbo {

constexpr int kTileM = 128;          // candidates per tile (TMEM lanes)
constexpr int kSlabK = 64;           // features per slab (= one 128-byte swizzle row of fp16)
constexpr int kUmmaK = 16;           // K per tcgen05.mma for 16-bit inputs
// epilogue + producers + TMA + 2 projection + 2 main MMA warps
constexpr int kProjWarpsFor(int npg) { return 2; }
constexpr int kNumThreadsFor(int npg) { return 32 * (4 + 4 * npg + 1 + kProjWarpsFor(npg) + 2); }
// CTA pairs add a second set of four epilogue warps (one per TMEM lane quarter): the accumulator read-out is latency-bound
// with one warp per quarter, and at SC = 256 it is not overlapped (single accumulator buffer)
constexpr int kNumThreadsPair(int npg) { return kNumThreadsFor(npg) + 128; }
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 12;       // smem ring (Theta + W tiles)
constexpr int kMaxWStages = 8;     // separate (W | b) tile ring (split mode)
constexpr int kMaxRing = 6;          // TMEM ring (projection result -> Phi hi/lo), 64 columns per slot
constexpr int kCopelandFixBits = 23;                     // soft-Copeland sums: sigmoid in 2^-23 fixed point
constexpr float kCopelandFixScale = 8388608.f;           // 2^23, also the magic number that rounds s 2^23 to an integer

// Projection K extent for a template input dimension: 6 split products per dimension + 3 bias terms.
__host__ __device__ constexpr int proj_k_used(int DT) { return ((6 * DT + 3 + 15) / 16) * 16; }     // K multiplied
__host__ __device__ constexpr int proj_k_alloc(int DT) {                                            // K stored (pow2)
  return proj_k_used(DT) <= 16 ? 16 : (proj_k_used(DT) <= 32 ? 32 : (proj_k_used(DT) <= 64 ? 64 : 128));
}
__host__ __device__ constexpr int proj_row_bytes(int DT) { return (proj_k_alloc(DT) < 64 ? proj_k_alloc(DT) : 64) * 2; }
__host__ __device__ constexpr int proj_tiles(int DT) { return proj_k_alloc(DT) > 64 ? 2 : 1; }      // 64-wide K blocks
__host__ __device__ constexpr int proj_w_bytes(int DT) { return kSlabK * proj_row_bytes(DT) * proj_tiles(DT); }
// Bytes per smem stage: Theta hi + lo tiles (SC rows x 128 B each) + the projection tile(s), 1024-byte aligned.
__host__ __device__ constexpr int stage_tx_bytes(int SC, int DT) { return 2 * SC * 128 + proj_w_bytes(DT); }
__host__ __device__ constexpr int stage_stride_bytes(int SC, int DT) {
  return ((stage_tx_bytes(SC, DT) + 1023) / 1024) * 1024;
}

// CTA pairs (PAIR, cta_group::2): every CTA holds HALF of a stage's Theta rows (SC / 2) and half of the slab's (W | b)
// rows (32 features); the pair's MMAs of M = 256 read both halves.
__host__ __device__ constexpr int pair_stage_tx_bytes(int SC, int DT) { return SC * 128 + proj_w_bytes(DT) / 2; }
__host__ __device__ constexpr int pair_stage_stride_bytes(int SC, int DT) {
  return ((pair_stage_tx_bytes(SC, DT) + 1023) / 1024) * 1024;
}

// With four producer groups (small sample chunks, where the SFU binds) and with CTA pairs the split candidate rows live in shared memory
// instead of TMEM, which frees the columns of a sixth ring slot: 128 rows in the same K-major swizzled tile format as
// the (W | b) tile, two buffers after the stage ring; the projection MMA then takes both operands from shared memory.
__host__ __device__ constexpr int proj_x_bytes(int DT) { return kTileM * proj_row_bytes(DT) * proj_tiles(DT); }

struct RffParams {
  const float* X;            // [N, d] candidates
  const float* unscale;      // [Salloc] 1 / (power-of-two scale folded into Theta)
  unsigned long long* packed;  // [Salloc] running (orderable value << 32 | ~row); zero between launches
  unsigned int* done_counter;  // last-CTA detection
  float* out_val;            // [S]
  long long* out_idx;        // [S]
  long long* out_key;        // [S] or null: signed-orderable (value, ~global index) key instead of out_val/out_idx
  // multi-GPU merge without a collective call: the last CTA stores this rank's S keys into slot `peer_rank` of EVERY
  // rank's peer buffer (NVLink peer stores; buffers mapped into this process, layout in rff.cu: peer_buffer_*) and then
  // releases flag[peer_rank] = peer_epoch there; each rank's merge kernel waits for its own flags.  peer_world = 0: off.
  unsigned long long* peer_buf[8];
  int peer_world, peer_rank;
  unsigned long long peer_epoch;
  // peer_merge = 1: the last CTA also does the merge -- after its own delivery it acquires the `peer_world` flags of its
  // OWN buffer (peer_buf[peer_rank]), takes the per-sample maximum over the slots and writes the global (value, index) to
  // out_val / out_idx: one launch per rank and iteration.  A peer that does not deliver within ~10 s gives index -2 for
  // every sample and the epoch in *peer_status (host-mapped word, may be null).
  int peer_merge;
  unsigned int* peer_status;
  // streamed candidates (tkbo_rff_argmax_host): X arrives in chunks of `rows_per_chunk` rows by H2D copies on another
  // stream while this launch runs; ready_flag[0] = number of chunks landed so far (written by a copy queued behind each
  // chunk), ready_flag[1] = set by the kernel if a chunk never arrived.  Null: all of X is resident.
  unsigned int* ready_flag;
  long long rows_per_chunk;
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
  int stages;                // smem ring depth
  int ring;                  // TMEM ring depth (64-column slots)
  int acc_bufs;              // 1 or 2 accumulator buffers of SC columns
  int xa_bufs;               // 1 or 2 buffers for the split candidate rows
  int proj_warps;            // 1 or 2 projection-issuing warps (2 needs an even smem ring depth)
  int coop;                  // 1 (two producer groups only): both groups convert every slab, 32 columns each -- half the
                             // per-slab latency, for shapes whose TMEM ring has only 3 slots
  int wstages;               // 0: the (W | b) tile of a slab travels in its Theta stage; > 0: it has its own ring of this
                             // depth, so projections + feature producers run ahead of the (few, large) Theta stages
  int pair;                  // 1: launched as clusters of two CTAs that share every tcgen05.mma (template PAIR)
  int halves;                // 1 (pairs, one accumulator buffer, SC % 64 == 0): the accumulator is retired in two column halves,
                             // so the next item's first MMAs overlap the read-out of the second half
  // Two-pass argmax (MODE 0).  Pass 1 (approx = 1) runs the hi*hi MMAs only - half (fast form) or a third (fp32 form) of the
  // tensor work.  Per sample it keeps (largest approximate value seen) - 2 eps_s in the running keys (eps2[s] = 2 eps_s,
  // eps_s >= |full - approximate| in accumulator units) and stores the approximate maximum of every group of 32 rows in
  // rg_max; its last CTA hands the final bounds out as lb_out.  mark_items_kernel then lists the items in which some row
  // group comes within 2 eps of the largest approximate value - everything else provably cannot hold (or tie) the
  // maximum - and pass 2 is the ordinary launch restricted to that list (item_list != nullptr).
  int approx;
  const float* eps2;         // [Salloc]
  float* rg_max;             // pass 1: [ceil(N / 32)][Salloc] approximate maxima per row group (accumulator units)
  float* lb_out;             // pass 1: [Salloc] final lower bounds
  const int* item_list;      // pass 2
  const int* list_count;     // pass 2
  int lead;                  // halves mode: slabs whose half-0 MMAs are issued ahead at the start of an item (0 = the ring depth)
  int nsplit;                // single CTAs: 1 = the two main MMA warps each issue half of the sample columns (N = SC / 2) of
                             // every slab; 0 = one warp issues N = SC.  An MMA with A in TMEM costs >= ~32 cycles
                             // whatever its N (tools/trace_k1.py: 12 MMAs of N = 32 take ~480 cycles to issue), so
                             // chunks of <= 128 samples must not be split: the MMA count per slab would bind
  int duel_n;                // > 0: the candidates are the ordered pairs [x_i, x_j] of duel_n base points of dimension d / 2
                             // held in X (dts.py:31-45); a tile = one i and 128 consecutive j; never materialised
  int duel_j0, duel_nj;      // this launch's shard of the second argument: j in [duel_j0, duel_j0 + duel_nj)
  unsigned long long* score_fix;   // [S, duel_n] (MODE 2) sum over j of sigmoid(f) in 2^-23 fixed point (deterministic atomics)
  int debug;                 // timing ablations (TKBO_DEBUG): 1 = hi*hi MMA only, 2 = producers skip the cos, 4 = no Theta TMA
  unsigned long long* trace; // optional clock64() event log of CTA 0 ([kTraceEvents][trace_cap]), tools/trace_k1.py
  int trace_cap;
};

// Pipeline trace events (CTA 0, lane 0 of one warp per role); index = that role's own iteration counter.
enum TraceEvent : int {
  kTrTmaEmpty = 0, kTrTmaIssued = 1,
  kTrProdPFull = 2 /* + 2 * group */, kTrProdArrived = 3 /* + 2 * group */,
  kTrMmaProj = 10, kTrMmaAFull = 11, kTrMmaCommitted = 12,
  kTrEpiAccFull = 13, kTrEpiDone = 14, kTrMmaAccEmpty = 15, kTrMmaProjIssued = 16, kTraceEvents = 17
};
// Compiled in only with -DTKBO_TRACE (libtkbo_trace.so): the issuing warps' loops must stay as short as possible.
__device__ __forceinline__ void trace_event(const RffParams& p, int ev, int idx) {
#ifdef TKBO_TRACE
  if (p.trace != nullptr && blockIdx.x == 0 && (threadIdx.x & 31) == 0 && idx < p.trace_cap)
    p.trace[ev * p.trace_cap + idx] = static_cast<unsigned long long>(clock64());
#endif
}

// u - h for a packed pair of halves h (mixed-precision fma: h * -1 + u, one instruction per value, exact)
__device__ __forceinline__ void sub_half2(float u0, float u1, uint32_t h2, float& l0, float& l1) {
  asm("{\n\t.reg .b16 a, b, m;\n\t"
      "mov.b32 {a, b}, %2;\n\t"
      "mov.b16 m, 0xBC00;\n\t"
      "fma.rn.f32.f16 %0, a, m, %3;\n\t"
      "fma.rn.f32.f16 %1, b, m, %4;\n\t}"
      : "=f"(l0), "=f"(l1) : "r"(h2), "f"(u0), "f"(u1));
}
// The producers' cos; the timing ablation that skips it (TKBO_DEBUG & 2) exists in the trace build only.
__device__ __forceinline__ float feature_cos(uint32_t arg_bits, bool skip) {
#ifdef TKBO_TRACE
  if (skip) return 1.f;
#endif
  return __cosf(__uint_as_float(arg_bits));
}

// cos on the FMA pipe: Cody-Waite reduction to [-pi, pi] (magic-number rint, 2 pi in two parts) and a degree-6 minimax
// polynomial in r^2 (max error 3.3e-7, MUFU.COS: ~5e-7): 11 FMA-pipe instructions and no XU slot.  Meant for the shapes
// the feature map binds (sample chunks <= 64: XU 74 % busy, FMA 25 %): a share of the features of every slab would take
// this route so that both pipes fill up; kPolyMask = which of 16 consecutive features (bit e = element e).
// MEASURED (B200, c2, same box): 0/16 features 0.369 ms, 4/16 0.420, 5/16 0.438, 6/16 0.452, 8/16 0.487 -- the
// producers are bound by issue slots / dependent-instruction latency, not by the XU pipe alone, and every polynomial
// costs nine instructions more than FMUL + MUFU.  Kept compiled out (mask 0) as a documented negative result.
#ifndef TKBO_POLY_MASK
#define TKBO_POLY_MASK 0x0u         // off: measured slower, see the comment above (0x4924u = elements 2, 5, 8, 11, 14)
#endif
constexpr uint32_t kPolyMask = TKBO_POLY_MASK;
__device__ __forceinline__ float poly_cos(float x) {
  const float nm = fmaf(x, 0.15915494309189535f, 12582912.f);      // + 1.5 * 2^23: the integer nearest to x / 2 pi
  const float n = nm - 12582912.f;
  float r = fmaf(n, -6.2831854820251465f, x);                        // 2 pi = 6.2831854820251465 - 1.7484555e-7
  r = fmaf(n, 1.7484555e-7f, r);
  const float u = r * r;
  float p = fmaf(u, 1.724442811e-09f, -2.707885187e-07f);
  p = fmaf(u, p, 2.476986447e-05f);
  p = fmaf(u, p, -1.388780307e-03f);
  p = fmaf(u, p, 4.166648909e-02f);
  p = fmaf(u, p, -4.999998808e-01f);
  return fmaf(u, p, 1.f);
}

// (Tried: the residual as a truncated bf16 -- one PRMT instead of an F2FP -- multiplied against the fp16 Theta tile.  A
// kind::f16 MMA with A = bf16 and B = fp16 is an illegal instruction on sm_100a; and F2FP does not compete with MUFU
// anyway: tools/probe_alu_rates.cu measures COS + F2FP at the rate of COS alone.)
// 16 projection values (fp32 bits) -> 8 packed fp16x2 hi words followed by 8 lo words: five instructions per feature
// (FMUL.RZ + MUFU.COS, half an F2FP for hi, one FHFMA for the residual, half an F2FP for lo).  POLY: the features in
// kPolyMask evaluate their cos on the FMA pipe instead.
template <bool POLY = false>
__device__ __forceinline__ void cos_split16(const uint32_t* __restrict__ arg, uint32_t (&out)[16], bool skip) {
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    const float c0 = (POLY && ((kPolyMask >> (2 * jj)) & 1u)) ? poly_cos(__uint_as_float(arg[2 * jj])) : feature_cos(arg[2 * jj], skip);
    const float c1 = (POLY && ((kPolyMask >> (2 * jj + 1)) & 1u)) ? poly_cos(__uint_as_float(arg[2 * jj + 1]))
                                                                  : feature_cos(arg[2 * jj + 1], skip);
    const __half2 h = __floats2half2_rn(c0, c1);
    const uint32_t hw = *reinterpret_cast<const uint32_t*>(&h);
    float l0, l1;
    sub_half2(c0, c1, hw, l0, l1);
    const __half2 l = __floats2half2_rn(l0, l1);
    out[jj] = hw;
    out[8 + jj] = *reinterpret_cast<const uint32_t*>(&l);
  }
}

// Fast-precision variant (C8): 16 projection values -> 8 fp16x2 words of u = 2^12 cos (the hi operand; the 2^-12 is folded
// into the per-sample unscale), then the operand of ONE e4m3 MMA that carries both correction terms: 4 words e4m3(cos)
// (against e4m3(2^12 Theta_lo)) and 4 words e4m3(u - fp16(u)) (against e4m3(Theta)); byte order = feature order.
// ~3e-5 of max|f| instead of ~2e-6, 8 instead of 12 MMAs per slab, six instructions per feature.
__device__ __forceinline__ uint32_t pack_e4m3x4(float a0, float a1, float a2, float a3) {
  uint16_t lo, hi;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(a1), "f"(a0));      // first source -> upper byte
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(a3), "f"(a2));
  return static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
}
__device__ __forceinline__ void cos_split16_c8(const uint32_t* __restrict__ arg, uint32_t (&out)[16], bool skip) {
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    float c[4], u[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      c[e] = feature_cos(arg[4 * jj + e], skip);
      u[e] = c[e] * 4096.f;
    }
    const __half2 h01 = __floats2half2_rn(u[0], u[1]);
    const __half2 h23 = __floats2half2_rn(u[2], u[3]);
    const uint32_t w01 = *reinterpret_cast<const uint32_t*>(&h01), w23 = *reinterpret_cast<const uint32_t*>(&h23);
    sub_half2(u[0], u[1], w01, l[0], l[1]);
    sub_half2(u[2], u[3], w23, l[2], l[3]);
    out[2 * jj] = w01;
    out[2 * jj + 1] = w23;
    out[8 + jj] = pack_e4m3x4(c[0], c[1], c[2], c[3]);
    out[12 + jj] = pack_e4m3x4(l[0], l[1], l[2], l[3]);
  }
}

// Pass 1 of the two-pass argmax: only the hi operand is read, so the residual / e4m3 words are left as zeros (three to
// three and a half instructions per feature).
template <bool C8v>
__device__ __forceinline__ void cos_hi16(const uint32_t* __restrict__ arg, uint32_t (&out)[16], bool skip) {
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {
    float c0 = feature_cos(arg[2 * jj], skip), c1 = feature_cos(arg[2 * jj + 1], skip);
    if constexpr (C8v) { c0 *= 4096.f; c1 *= 4096.f; }
    const __half2 h = __floats2half2_rn(c0, c1);
    out[jj] = *reinterpret_cast<const uint32_t*>(&h);
    out[8 + jj] = 0u;
  }
}

// x = x1 + x2 + x3 exactly, each term a bf16 (truncating the fp32 bits keeps every step exact).
__device__ __forceinline__ void bf16_split3(float x, uint32_t& b1, uint32_t& b2, uint32_t& b3) {
  const uint32_t u = __float_as_uint(x);
  const float r1 = x - __uint_as_float(u & 0xFFFF0000u);
  const uint32_t v = __float_as_uint(r1);
  const float r2 = r1 - __uint_as_float(v & 0xFFFF0000u);      // <= 8 significant bits: exactly a bf16
  b1 = u >> 16;
  b2 = v >> 16;
  b3 = __float_as_uint(r2) >> 16;
}

// PAIR (clusters of two CTAs, cta_group::2): a work item is TWO adjacent 128-candidate tiles x one sample chunk.  Each CTA
// keeps its own rows end to end (staging, producers, epilogue, TMEM), but every tcgen05.mma is issued once, by the leader
// CTA, with M = 256: the B operands (Theta tiles, (W | b) tile) are fetched half by each CTA, so the L2 -> shared-memory
// traffic and the shared-memory reads per SM halve and a stage is half as large (twice as many fit).  Cross-CTA hand-offs:
// TMA of both CTAs completes on the leader's b_full; the peer's producer / staging / epilogue warps arrive on the leader's
// a_full / xa_full / acc_empty through the cluster address space; the leader's commits are multicast to both CTAs'
// p_full / b_empty / acc_full.
template <int DT, int MODE /* 0 = argmax, 1 = store F, 2 = duel soft-Copeland sums */, int NPG /* producer warp groups: 2, 3 or 4 */,
          bool C8 /* corrections as one e4m3 MMA (tm_lo = the e4m3 tile) instead of two fp16 MMAs */, bool PAIR = false>
__global__ void __launch_bounds__(PAIR ? kNumThreadsPair(NPG) : kNumThreadsFor(NPG), 1)
rff_fused_kernel(const __grid_constant__ CUtensorMap tm_hi, const __grid_constant__ CUtensorMap tm_lo,
                 const __grid_constant__ CUtensorMap tm_w, const RffParams p) {
  constexpr int kPUsed = proj_k_used(DT);
  constexpr int kPCols = proj_k_alloc(DT) / 2;       // 32-bit words of one split candidate row = TMEM columns of one buffer
  constexpr bool kXSmem = (NPG == 4) || PAIR;        // split candidate rows in shared memory (else in TMEM)
  constexpr int kXBytes = proj_x_bytes(DT);
  constexpr int kPRow = proj_row_bytes(DT);
  constexpr int kPTile = (PAIR ? kSlabK / 2 : kSlabK) * kPRow;     // bytes of one 64-wide K block of the (W | b) rows held here
  extern __shared__ uint8_t smem_raw[];

  // mbarriers, addressed as bar0 + 8 * index in the hot loops
  enum : int {
    kBFull = 0,                        // [kMaxStages] TMA landed the Theta + W tiles of a slab
    kBEmpty = kBFull + kMaxStages,     // [kMaxStages] main MMAs reading the stage finished
    kPFull = kBEmpty + kMaxStages,     // [kMaxRing] projection of a slab is in its TMEM ring slot
    kAFull = kPFull + kMaxRing,        // [kMaxRing] producers replaced it by the Phi hi/lo operand
    kAccFull = kAFull + kMaxRing,      // [2]
    kAccEmpty = kAccFull + 2,          // [2]
    kXaFull = kAccEmpty + 2,           // [2] split candidate rows of an item are in TMEM
    kWFull = kXaFull + 2,              // [kMaxWStages] split mode: TMA landed the (W | b) tile of a slab
    kWEmpty = kWFull + kMaxWStages,    // [kMaxWStages] its projection MMAs finished
    kNumBars = kWEmpty + kMaxWStages
  };
  __shared__ uint64_t bars[kNumBars];
  __shared__ uint32_t tmem_base_smem;
  __shared__ int is_last_cta;
  __shared__ int peer_timed_out;
  __shared__ __align__(16) float thr_s[256];       // per column of the current chunk: running maxima (MODE 0) / unscale (MODE 1, 2)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // kernel parameters used in the loops, as registers
  const int SC = p.SC, n_slabs = p.n_slabs, n_chunks = p.n_chunks, stages = p.stages, ring = p.ring;
  const int acc_bufs = p.acc_bufs, xa_bufs = p.xa_bufs;
  const int wstages = p.wstages;
  const bool split_w = wstages > 0;
  const int sbytes = PAIR ? pair_stage_stride_bytes(SC, DT) : (split_w ? 2 * SC * 128 : stage_stride_bytes(SC, DT));
  const uint32_t cta = PAIR ? cluster_ctarank() : 0u;              // 0 = leader of the pair
  const bool leader = (cta == 0u);
  const int unit = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);      // item stream of this CTA / pair
  const int n_units = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int n_items_all = (PAIR ? (p.n_mtiles + 1) / 2 : p.n_mtiles) * n_chunks;
  // The two-pass argmax exists for the tensor-bound forms only (the host never asks the four-group form for it): compiled out
  // of the NPG = 4 instances, whose timing is sensitive to every instruction around their producers (c2: 0.365 -> 0.43 ms
  // with the run-time branches in place)
  constexpr bool kTwoPass = (MODE == 0) && (NPG != 4);
  const bool approx_pass = kTwoPass && p.approx != 0;
  const bool listed = kTwoPass && (p.item_list != nullptr) && !approx_pass;   // pass 2: items come from the list
  const int n_items = listed ? __ldcg(p.list_count) : n_items_all;
  auto item_of = [&](int i) { return listed ? __ldcg(p.item_list + i) : i; };
  const int my_items = (n_items - unit + n_units - 1) / n_units;
  const int th_rows = PAIR ? SC / 2 : SC;                          // Theta rows of a stage held by this CTA
  const int total = my_items * n_slabs;              // slab iterations of this CTA
  const uint32_t bar0 = smem_u32(bars);
  // 1024-byte alignment for the 128B-swizzled tiles
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&bars[kBFull + s], 1);
      mbar_init(&bars[kBEmpty + s], (PAIR || !p.nsplit) ? 1 : 2);      // one commit per issuing main MMA warp
    }
    for (int s = 0; s < kMaxRing; ++s) {
      mbar_init(&bars[kPFull + s], 1);
      mbar_init(&bars[kAFull + s], (p.coop ? 8 : 4) * (PAIR ? 2 : 1));   // one arrival per producer warp working on the slab
    }
    for (int s = 0; s < kMaxWStages; ++s) {
      mbar_init(&bars[kWFull + s], 1);
      mbar_init(&bars[kWEmpty + s], 1);      // one commit by the projection warp that used it
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars[kAccFull + a], (PAIR || !p.nsplit) ? 1 : 2);     // one commit per issuing main MMA warp
      mbar_init(&bars[kAccEmpty + a], PAIR ? 16 : 4);   // one arrival per epilogue warp (PAIR: two sets of four, both CTAs)
      mbar_init(&bars[kXaFull + a], PAIR ? 8 : 4);
    }
    fence_mbar_init();
  }
  constexpr int kTmaWarp = 4 + 4 * NPG;
  constexpr int kProjWarp = kTmaWarp + 1;            // kProjWarps of them
  constexpr int kProjWarps = kProjWarpsFor(NPG);
  constexpr int kMainWarp = kProjWarp + kProjWarps;  // and + 1
  constexpr int kEpiSets = PAIR ? 2 : 1;             // PAIR: warps kMainWarp + 2 .. + 5 are a second set of epilogue warps
  if (warp == kMainWarp) {
    if constexpr (PAIR) tmem_alloc_pair(&tmem_base_smem, kTmemCols); else tmem_alloc(&tmem_base_smem, kTmemCols);
  }
  if (warp == kTmaWarp && lane == 0) {
    tma_prefetch_desc(&tm_hi);
    tma_prefetch_desc(&tm_lo);
    tma_prefetch_desc(&tm_w);
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();    // barriers initialised (cluster-wide) + TMEM allocated
  tc_fence_after();
  // barriers of the leader as seen from either CTA (the peer arrives there); in the leader these are its own
  const uint32_t bar0_lead = PAIR ? mapa_shared(bar0, 0u) : bar0;
  const uint32_t tmem_base = tmem_base_smem;
  const uint32_t xa_col0 = acc_bufs * SC;                         // split candidate rows follow the accumulators (TMEM form)
  const uint32_t w_smem0 = smem0 + stages * sbytes;               // split mode: the (W | b) ring follows the Theta stages (never with PAIR)
  const uint32_t xa_smem0 = w_smem0 + wstages * proj_w_bytes(DT); // candidate rows, smem form: two buffers after the rings
  const uint32_t ring_col0 = kTmemCols - ring * kSlabK;           // the ring occupies the top columns

  if (warp == kTmaWarp) {
    // ------------------------------------------------------------------ TMA producer (warp-converged, one elected issuer)
    int st = 0, it = 0;
    uint32_t ph = 0;
    const bool theta = !(p.debug & 4);               // ablation 4: no Theta tile loads
    const uint32_t txbytes = split_w ? 2 * SC * 128 : ((p.debug & 4) ? proj_w_bytes(DT) : stage_tx_bytes(SC, DT));
    // split mode: (W | b) tiles run up to `wstages` slabs ahead of the Theta tiles; a tile beyond the current slab is only
    // fetched if its stage is already free (test), the current slab's tile is waited for
    int jw = 0, wst = 0;
    uint32_t wph = 0;
    auto issue_w = [&](int slab_w, uint32_t dst, uint32_t full) {
#pragma unroll
      for (int t = 0; t < proj_tiles(DT); ++t) tma_load_2d(dst + t * kPTile, &tm_w, full, t * 64, slab_w * kSlabK);
    };
    for (int it_i = unit; it_i < n_items; it_i += n_units) {
      const int item = item_of(it_i);
      const int chunk = item % n_chunks;
      for (int slab = 0; slab < n_slabs; ++slab, ++it) {
        if constexpr (PAIR) {
          // both CTAs: wait for the stage to be free here (the leader's commit is multicast), fetch this CTA's half of the
          // Theta rows and of the (W | b) rows; every byte of the pair completes on the LEADER's b_full
          mbar_wait(bar0 + 8 * (kBEmpty + st), ph ^ 1);
          if (elect_one()) {
            const uint32_t sb = smem0 + st * sbytes;
            const uint32_t full_c = bar0_lead + 8 * (kBFull + st);
            if (leader) mbar_arrive_expect_tx(bar0 + 8 * (kBFull + st), 2u * pair_stage_tx_bytes(SC, DT));
            if (p.halves) {
              // column half h of the accumulator = samples [h SC/2, (h+1) SC/2) of the chunk; an MMA of N = SC / 2 takes
              // SC / 4 of its B rows from each CTA, so this CTA keeps rows [h SC/2 + cta SC/4, + SC/4) at offset h SC/4
              const int qr = SC / 4;
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const int row0 = chunk * SC + hh * (SC / 2) + static_cast<int>(cta) * qr;
                tma_load_2d_pair(sb + hh * qr * 128, &tm_hi, full_c, slab * kSlabK, row0);
                tma_load_2d_pair(sb + th_rows * 128 + hh * qr * 128, &tm_lo, full_c, slab * kSlabK * (C8 ? 2 : 1), row0);
              }
            } else {
              const int row0 = chunk * SC + static_cast<int>(cta) * th_rows;
              tma_load_2d_pair(sb, &tm_hi, full_c, slab * kSlabK, row0);
              tma_load_2d_pair(sb + th_rows * 128, &tm_lo, full_c, slab * kSlabK * (C8 ? 2 : 1), row0);
            }
#pragma unroll
            for (int t = 0; t < proj_tiles(DT); ++t)
              tma_load_2d_pair(sb + 2 * th_rows * 128 + t * kPTile, &tm_w, full_c, t * 64,
                               slab * kSlabK + static_cast<int>(cta) * (kSlabK / 2));
          }
          __syncwarp();
          if (++st == stages) { st = 0; ph ^= 1; }
          continue;
        }
        if (split_w) {
          while (jw < total && jw < it + wstages) {
            const uint32_t emp = bar0 + 8 * (kWEmpty + wst);
            if (jw <= it) mbar_wait(emp, wph ^ 1);
            else if (!mbar_test_wait(emp, wph ^ 1)) break;
            if (elect_one()) {
              const uint32_t full = bar0 + 8 * (kWFull + wst);
              mbar_arrive_expect_tx(full, proj_w_bytes(DT));
              issue_w(jw % n_slabs, w_smem0 + wst * proj_w_bytes(DT), full);
            }
            __syncwarp();
            ++jw;
            if (++wst == wstages) { wst = 0; wph ^= 1; }
          }
        }
        mbar_wait(bar0 + 8 * (kBEmpty + st), ph ^ 1);
        trace_event(p, kTrTmaEmpty, it);
        if (elect_one()) {
          const uint32_t sb = smem0 + st * sbytes;
          const uint32_t full = bar0 + 8 * (kBFull + st);
          if (split_w && !theta) mbar_arrive(full);
          else mbar_arrive_expect_tx(full, txbytes);
          if (theta) {
            tma_load_2d(sb, &tm_hi, full, slab * kSlabK, chunk * SC);
            tma_load_2d(sb + SC * 128, &tm_lo, full, slab * kSlabK * (C8 ? 2 : 1), chunk * SC);   // C8: byte tensor
          }
          if (!split_w) issue_w(slab, sb + 2 * SC * 128, full);
        }
        __syncwarp();
        trace_event(p, kTrTmaIssued, it);
        if (++st == stages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp >= kProjWarp && warp < kProjWarp + kProjWarps) {
    // ------------------------------------------------------------------ projection MMAs: arg = [x splits | 1] . [w splits | b]^T
    // Warp h of `npw` takes the slab iterations j = h, h + npw, ...  A ring slot is free again once the main MMAs of
    // slab j - ring have finished, which is b_empty of that slab's smem stage (stages >= ring, and b_full(j) seen
    // first implies every earlier phase of that barrier has completed).  The candidate rows of item i+2 (i+1 with
    // one buffer) are only staged after the epilogue of item i, i.e. after all of item i's projections were consumed.
    const int npw = PAIR ? 1 : (p.proj_warps < kProjWarps ? p.proj_warps : kProjWarps);
    const int h = warp - kProjWarp;
    if (h < npw && leader) {
      constexpr uint32_t idesc_proj = idesc_bf16_f32(PAIR ? 2 * kTileM : kTileM, kSlabK);
      constexpr uint32_t kWDescHi = smem_desc_hi(8 * kPRow, kPRow == 128 ? 2u : (kPRow == 64 ? 4u : 6u));
      // the (W | b) tile(s) follow Theta hi, lo inside the stage, or (split mode) sit in their own ring
      const int wn = split_w ? wstages : stages;
      const uint32_t desc_stage = static_cast<uint32_t>(split_w ? proj_w_bytes(DT) : sbytes) >> 4;
      const uint32_t desc_w0 = smem_desc_lo(split_w ? w_smem0 : smem0 + 2 * th_rows * 128);
      const uint32_t wfull0 = bar0 + 8 * (split_w ? kWFull : kBFull);
      int slab = h % n_slabs, item = h / n_slabs;      // position of j inside this CTA's item sequence
      int st = h % wn, slot = h % ring;
      uint32_t ph = (h / wn) & 1;
      int fst = 0;                                     // smem stage of slab j - ring (valid once j >= ring)
      uint32_t fph = 0;
      bool new_item = true;
      // One row buffer = one barrier whose phases are the items in order; a parity wait only tells adjacent phases
      // apart, so a warp that owns no slab of an item (n_slabs < npw) still observes that item's phase, but never
      // the phase of an item beyond its own last slab iteration (it would not arrive).
      if (xa_bufs == 1 && h < total)
        for (int i = 0; i < item; ++i) mbar_wait_cluster(bar0 + 8 * kXaFull, static_cast<uint32_t>(i) & 1);
      for (int j = h; j < total; j += npw) {
        mbar_wait(wfull0 + 8 * st, ph);
        if (j >= ring) mbar_wait(bar0 + 8 * (kBEmpty + fst), fph);
        if (new_item) {
          const int xb = (xa_bufs == 2) ? (item & 1) : 0;
          const uint32_t use = (xa_bufs == 2) ? static_cast<uint32_t>(item >> 1) : static_cast<uint32_t>(item);
          mbar_wait_cluster(bar0 + 8 * (kXaFull + xb), use & 1);
          new_item = false;
        }
        trace_event(p, kTrMmaProj, j);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t d_tmem = tmem_base + ring_col0 + slot * kSlabK;
          const int xb = (xa_bufs == 2) ? (item & 1) : 0;
          const uint32_t a_tmem = tmem_base + xa_col0 + xb * kPCols;
          const uint32_t a_lo = smem_desc_lo(xa_smem0 + xb * kXBytes);
          const uint32_t b_lo = desc_w0 + st * desc_stage;
#pragma unroll
          for (int k = 0; k < kPUsed / kUmmaK; ++k) {
            const uint32_t koff = (k % 4) * ((kUmmaK * 2) >> 4);
            const uint64_t b_desc = (static_cast<uint64_t>(kWDescHi) << 32) | (b_lo + (k / 4) * (kPTile >> 4) + koff);
            if constexpr (PAIR) {
              mma_f16_ss_pair(d_tmem, (static_cast<uint64_t>(kWDescHi) << 32) | (a_lo + (k / 4) * ((kTileM * kPRow) >> 4) + koff),
                              b_desc, idesc_proj, k != 0);
            } else if constexpr (kXSmem) {
              mma_f16_ss(d_tmem, (static_cast<uint64_t>(kWDescHi) << 32) | (a_lo + (k / 4) * ((kTileM * kPRow) >> 4) + koff),
                         b_desc, idesc_proj, k != 0);
            } else {
              mma_f16_ts(d_tmem, a_tmem + k * (kUmmaK / 2), b_desc, idesc_proj, k != 0);
            }
          }
          if constexpr (PAIR) tc_commit_pair(bar0 + 8 * (kPFull + slot), 3);
          else tc_commit(bar0 + 8 * (kPFull + slot));
          if (split_w) tc_commit(bar0 + 8 * (kWEmpty + st));
        }
        __syncwarp();
        trace_event(p, kTrMmaProjIssued, j);
        // advance by npw slab iterations
        if (j >= ring) { fst += npw; if (fst >= stages) { fst -= stages; fph ^= 1; } }
        else if (j + npw >= ring) { fst = (j + npw - ring) % stages; fph = 0; }
        st += npw; if (st >= wn) { st -= wn; ph ^= 1; }
        slot += npw; if (slot >= ring) slot -= ring;
        slab += npw;
        while (slab >= n_slabs) {
          slab -= n_slabs;
          ++item;
          new_item = true;
          if (slab >= n_slabs && xa_bufs == 1 && j + npw < total) {   // an item without a slab of this warp
            mbar_wait_cluster(bar0 + 8 * kXaFull, static_cast<uint32_t>(item) & 1);
          }
        }
      }
    }
  } else if (warp >= kMainWarp && warp < kMainWarp + 2) {
    // ------------------------------------------------------------------ main MMAs: acc += Phi_slab . Theta_slab^T (3 per K step)
    // Warp h owns accumulator columns [h SC/2, (h+1) SC/2): N = SC/2 per instruction, Theta rows h SC/2 ... of the tile,
    // so every column's MMAs come from one thread in slab order (deterministic fp32 accumulation; MMAs issued by
    // different threads are NOT ordered against each other, even across an mbarrier hand-off).  A warp needs ~300
    // cycles of waits / commits / address arithmetic per burst, so when the next slab's operand is already there it
    // issues two slabs per burst; warp 0 pairs (0,1)(2,3).., warp 1 pairs (1,2)(3,4).., which keeps the two warps'
    // bookkeeping phases apart instead of idling the tensor pipe together.
    // PAIR: one warp of the leader issues for both CTAs, full N = SC per instruction (each CTA's smem holds SC / 2 rows)
    const int h = warp - kMainWarp;
    const bool whole = PAIR || !p.nsplit;            // this warp issues all SC columns (and the other main warp none)
    const int NH = whole ? SC : (SC >> 1);
    const uint32_t idesc = idesc_f16_f32(PAIR ? 2 * kTileM : kTileM, NH);
    const uint32_t idesc8 = idesc_e4m3_f32(PAIR ? 2 * kTileM : kTileM, NH);
    // descriptor low words advance by plain adds: (addr >> 4) never carries out of its 14-bit field
    const uint32_t desc_lo0 = smem_desc_lo(smem0 + (whole ? 0 : h * NH * 128));
    const uint32_t desc_stage = static_cast<uint32_t>(sbytes) >> 4;
    const uint32_t desc_lopart = static_cast<uint32_t>(th_rows * 128) >> 4;     // Theta-lo tile follows the hi tile
    const bool hi_only = (p.debug & 1) != 0 || approx_pass;
    const bool pairing = ring >= 5 && !(p.debug & 8);  // a pair holds two ring slots while its MMAs run
    const bool halves = PAIR && p.halves;
    const int total_main = ((PAIR && (h != 0 || !leader || halves)) || (!PAIR && whole && h != 0)) ? 0 : total;
    if constexpr (PAIR) {
      if (halves && h == 0 && leader) {
        // One accumulator buffer retired in two column halves.  Per item: the half-0 MMAs of the first `lead` slabs go out
        // as soon as the epilogue has drained half 0 of the previous item, i.e. while it is still reading half 1; then,
        // once half 1 is free, their half-1 MMAs and the stage releases; the remaining slabs issue both halves in turn.
        // Every column still sees its MMAs in slab order (deterministic accumulation).
        const int NQ = SC >> 1;
        const uint32_t idh = idesc_f16_f32(2 * kTileM, NQ);
        const uint32_t idh8 = idesc_e4m3_f32(2 * kTileM, NQ);
        const uint32_t b_half = static_cast<uint32_t>((SC / 4) * 128) >> 4;     // rows of half 1 inside a Theta tile
        auto issue_half = [&](int st_, int slot_, int sl, int hh) {
          const uint32_t d_tmem = tmem_base + hh * NQ;
          const uint32_t b_hi = desc_lo0 + st_ * desc_stage + hh * b_half;
          const uint32_t b_lo = b_hi + desc_lopart;
          const uint32_t a_st = tmem_base + ring_col0 + slot_ * kSlabK;
#pragma unroll
          for (int k = 0; k < kSlabK / kUmmaK; ++k) {
            const uint32_t koff = k * ((kUmmaK * 2) >> 4);
            const uint32_t a_hi = a_st + k * kUmmaK;
            const uint32_t a_lo = a_hi + kUmmaK / 2;
            mma_f16_ts_pair(d_tmem, a_hi, smem_desc_from_lo(b_hi + koff), idh, (sl | k) != 0);
            if (!hi_only) {
              if constexpr (C8) {
                mma_f8_ts_pair(d_tmem, a_lo, smem_desc_from_lo(b_lo + koff), idh8, 1);
              } else {
                mma_f16_ts_pair(d_tmem, a_hi, smem_desc_from_lo(b_lo + koff), idh, 1);
                mma_f16_ts_pair(d_tmem, a_lo, smem_desc_from_lo(b_hi + koff), idh, 1);
              }
            }
          }
        };
        int st = 0, slot = 0;
        uint32_t sph = 0, eph = 0;
        const int lead = min(min(ring, n_slabs), p.lead > 0 ? p.lead : ring);
        for (int item = 0; item < my_items; ++item) {
          // phase A: half 0 of the first `lead` slabs (their ring slots and stages stay held)
          mbar_wait_cluster(bar0 + 8 * (kAccEmpty + 0), eph ^ 1);
          {
            int st_a = st, slot_a = slot;
            uint32_t sph_a = sph;
            for (int t = 0; t < lead; ++t) {
              mbar_wait_cluster(bar0 + 8 * (kAFull + slot_a), sph_a);
              tc_fence_after();
              if (elect_one()) issue_half(st_a, slot_a, t, 0);
              __syncwarp();
              if (++st_a == stages) st_a = 0;
              if (++slot_a == ring) { slot_a = 0; sph_a ^= 1; }
            }
          }
          // phase B: their half-1 MMAs, then the stages / slots are released; the middle slabs issue both halves in turn
          mbar_wait_cluster(bar0 + 8 * (kAccEmpty + 1), eph ^ 1);
          tc_fence_after();
          // the last `tail` slabs mirror the start: all their half-0 MMAs first (half 0 is then complete and the epilogue
          // starts on it), then their half-1 MMAs, which run while half 0 is read
          // (it pins every ring slot once more per item, which only pays when the item is long: D = 1024, 16 slabs, loses
          // 2.7 % with it, D = 4096 gains 1 %)
          const int tail = (p.halves >= 2 && n_slabs >= 12 * lead) ? lead : 0;
          for (int t = 0; t < n_slabs - tail; ++t) {
            if (t >= lead) {
              mbar_wait_cluster(bar0 + 8 * (kAFull + slot), sph);
              tc_fence_after();
            }
            if (elect_one()) {
              if (t >= lead) issue_half(st, slot, t, 0);
              issue_half(st, slot, t, 1);
              tc_commit_pair(bar0 + 8 * (kBEmpty + st), 3);
              if (t == n_slabs - 1) {
                tc_commit_pair(bar0 + 8 * (kAccFull + 0), 3);
                tc_commit_pair(bar0 + 8 * (kAccFull + 1), 3);
              }
            }
            __syncwarp();
            if (++st == stages) st = 0;
            if (++slot == ring) { slot = 0; sph ^= 1; }
          }
          if (tail > 0) {
            int st_c = st, slot_c = slot;
            uint32_t sph_c = sph;
            for (int t = n_slabs - tail; t < n_slabs; ++t) {
              mbar_wait_cluster(bar0 + 8 * (kAFull + slot_c), sph_c);
              tc_fence_after();
              if (elect_one()) {
                issue_half(st_c, slot_c, t, 0);
                if (t == n_slabs - 1) tc_commit_pair(bar0 + 8 * (kAccFull + 0), 3);
              }
              __syncwarp();
              if (++st_c == stages) st_c = 0;
              if (++slot_c == ring) { slot_c = 0; sph_c ^= 1; }
            }
            for (int t = n_slabs - tail; t < n_slabs; ++t) {
              if (elect_one()) {
                issue_half(st, slot, t, 1);
                tc_commit_pair(bar0 + 8 * (kBEmpty + st), 3);
                if (t == n_slabs - 1) tc_commit_pair(bar0 + 8 * (kAccFull + 1), 3);
              }
              __syncwarp();
              if (++st == stages) st = 0;
              if (++slot == ring) { slot = 0; sph ^= 1; }
            }
          }
          eph ^= 1;
        }
      }
    }
    int st = 0, slot = 0, ab = 0, slab = 0, nitem = 0;
    uint32_t sph = 0, aph = 0, bph = 0;
    // combined mode: a_full implies (through p_full and the projection warp's wait) that the stage's Theta tiles have
    // landed; split mode: the Theta stage has its own wait
    for (int it = 0; it < total_main;) {
      if (slab == 0) {
        if constexpr (PAIR) mbar_wait_cluster(bar0 + 8 * (kAccEmpty + ab), aph ^ 1);
        else mbar_wait(bar0 + 8 * (kAccEmpty + ab), aph ^ 1);
        if (h == 0) trace_event(p, kTrMmaAccEmpty, nitem);
      }
      if (split_w) mbar_wait(bar0 + 8 * (kBFull + st), bph);
      if constexpr (PAIR) mbar_wait_cluster(bar0 + 8 * (kAFull + slot), sph);
      else mbar_wait(bar0 + 8 * (kAFull + slot), sph);
      // second slab of the burst: same item, pair alignment of this warp, operand already produced.  (A burst pays the
      // warp's fixed cost per iteration once: at c2 single slabs cost 0.410 ms, bursts of two 0.367 ms.  Longer bursts were
      // tried in round 2 - as a run-time loop and as four predicated copies - and lost 0.04-0.05 ms: this hand-written
      // pair, whose second copy ptxas predicates and interleaves with the first, is the fastest form found.)
      int st2 = st + 1, slot2 = slot + 1;
      uint32_t sph2 = sph;
      if (st2 == stages) st2 = 0;
      if (slot2 == ring) { slot2 = 0; sph2 ^= 1; }
      const uint32_t bph2 = (st2 == 0) ? (bph ^ 1) : bph;
      const bool two = pairing && (whole || ((it & 1) == h)) && (slab + 1 < n_slabs) &&
                       (PAIR ? mbar_test_wait_cluster(bar0 + 8 * (kAFull + slot2), sph2) : mbar_test_wait(bar0 + 8 * (kAFull + slot2), sph2)) &&
                       (!split_w || mbar_test_wait(bar0 + 8 * (kBFull + st2), bph2));
      if (h == 0) trace_event(p, kTrMmaAFull, it);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t d_tmem = tmem_base + ab * SC + (whole ? 0 : h * NH);
        auto issue_slab = [&](int st_, int slot_, int sl) {
          const uint32_t b_hi = desc_lo0 + st_ * desc_stage;
          const uint32_t b_lo = b_hi + desc_lopart;
          const uint32_t a_st = tmem_base + ring_col0 + slot_ * kSlabK;
#pragma unroll
          for (int k = 0; k < kSlabK / kUmmaK; ++k) {
            const uint32_t koff = k * ((kUmmaK * 2) >> 4);               // 32 bytes of K per step, in 16-byte units
            const uint32_t a_hi = a_st + k * kUmmaK;                     // 16 columns per K step: 8 hi words, 8 lo words
            const uint32_t a_lo = a_hi + kUmmaK / 2;
            if constexpr (PAIR) {
              mma_f16_ts_pair(d_tmem, a_hi, smem_desc_from_lo(b_hi + koff), idesc, (sl | k) != 0);
              if (!hi_only) {
                if constexpr (C8) {
                  mma_f8_ts_pair(d_tmem, a_lo, smem_desc_from_lo(b_lo + koff), idesc8, 1);
                } else {
                  mma_f16_ts_pair(d_tmem, a_hi, smem_desc_from_lo(b_lo + koff), idesc, 1);
                  mma_f16_ts_pair(d_tmem, a_lo, smem_desc_from_lo(b_hi + koff), idesc, 1);
                }
              }
            } else {
              mma_f16_ts(d_tmem, a_hi, smem_desc_from_lo(b_hi + koff), idesc, (sl | k) != 0);
              if (!hi_only) {
                if constexpr (C8) {
                  // K = 32 e4m3: [cos (16) | u - fp16(u) (16)] . [2^12 Theta_lo (16) | Theta (16)], 32 bytes of the lo tile
                  mma_f8_ts(d_tmem, a_lo, smem_desc_from_lo(b_lo + koff), idesc8, 1);
                } else {
                  mma_f16_ts(d_tmem, a_hi, smem_desc_from_lo(b_lo + koff), idesc, 1);
                  mma_f16_ts(d_tmem, a_lo, smem_desc_from_lo(b_hi + koff), idesc, 1);
                }
              }
            }
          }
          if constexpr (PAIR) {
            tc_commit_pair(bar0 + 8 * (kBEmpty + st_), 3);
            if (sl == n_slabs - 1) tc_commit_pair(bar0 + 8 * (kAccFull + ab), 3);
          } else {
            tc_commit(bar0 + 8 * (kBEmpty + st_));
            if (sl == n_slabs - 1) tc_commit(bar0 + 8 * (kAccFull + ab));
          }
        };
        issue_slab(st, slot, slab);
        if (two) issue_slab(st2, slot2, slab + 1);
      }
      __syncwarp();
      if (h == 0) trace_event(p, kTrMmaCommitted, it);
      const int adv = two ? 2 : 1;
      it += adv;
      st += adv; if (st >= stages) { st -= stages; bph ^= 1; }
      slot += adv; if (slot >= ring) { slot -= ring; sph ^= 1; }
      slab += adv;
      if (slab == n_slabs) {
        slab = 0;
        ++nitem;
        if (++ab == acc_bufs) { ab = 0; aph ^= 1; }
      }
    }
  } else if (warp >= 4 && warp < kTmaWarp) {
    // ------------------------------------------------------------------ feature producers
    // group g owns the slab iterations it = g, g + NPG, g + 2 NPG, ... of this CTA, whatever item they belong to
    const int group = (warp - 4) >> 2;               // 0 .. NPG-1
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;
    const bool skip = (p.debug & 2) != 0;
    const bool approx = approx_pass;                 // pass 1 of the two-pass argmax reads the hi operand only
    auto split16 = [approx](const uint32_t* a, uint32_t (&o)[16], bool sk) {
      // four producer groups = the shapes bound by the feature map (XU pipe): part of the cos goes to the FMA pipe
      if (approx) cos_hi16<C8>(a, o, sk);
      else if constexpr (C8) cos_split16_c8(a, o, sk);
      else cos_split16<NPG == 4>(a, o, sk);
    };
    if ((NPG == 2 || NPG == 4) && p.coop) {
      // latency form: every slab is converted by two groups, each taking 32 of its 64 columns.  Two groups: both work on
      // every slab; four groups: the pair (0, 1) takes the even slab iterations, (2, 3) the odd ones -- a slab is then held
      // by its producers for half as long, which is what the TMEM ring (six slots of 64 columns) is short of at S <= 64
      constexpr int kStep = NPG / 2;                 // slab iterations between two slabs of this warp
      const int first_it = (NPG == 4) ? (group >> 1) : 0;
      int slot = first_it % ring;
      uint32_t ph = (first_it / ring) & 1;
      for (int it = first_it; it < total; it += kStep) {
        mbar_wait(bar0 + 8 * (kPFull + slot), ph);
        if (q == 0) trace_event(p, kTrProdPFull + 2 * group, it);
        tc_fence_after();
        const uint32_t col = tmem_base + lane_base + ring_col0 + slot * kSlabK + (group & 1) * 32;
        uint32_t r0[16], r1[16], out[16];
        tmem_ld_x16(col, r0);
        tmem_ld_x16(col + 16, r1);
        tmem_wait_ld();
        split16(r0, out, skip);
        tmem_st_x16(col, out);
        split16(r1, out, skip);
        tmem_st_x16(col + 16, out);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster_relaxed(bar0_lead + 8 * (kAFull + slot)); else mbar_arrive(bar0 + 8 * (kAFull + slot)); }
        if (q == 0) trace_event(p, kTrProdArrived + 2 * group, it);
        slot += kStep;
        if (slot >= ring) { slot -= ring; ph ^= 1; }
      }
    } else {
    int slot = group % ring;
    uint32_t ph = (group / ring) & 1;
    int n = 0;
    for (int it = group; it < total; it += NPG, ++n) {
      mbar_wait(bar0 + 8 * (kPFull + slot), ph);
      if (q == 0) trace_event(p, kTrProdPFull + 2 * group, n);
      tc_fence_after();
      const uint32_t col = tmem_base + lane_base + ring_col0 + slot * kSlabK;
      // 16 columns at a time, the next load in flight while the current values are converted (keeps the live
      // registers at two 16-word buffers + one output buffer)
      uint32_t r0[16], r1[16], out[16];
      tmem_ld_x16(col, r0);
      tmem_ld_x16(col + 16, r1);
      tmem_wait_ld();
      split16(r0, out, skip);
      tmem_st_x16(col, out);
      tmem_ld_x16(col + 32, r0);
      split16(r1, out, skip);
      tmem_st_x16(col + 16, out);
      tmem_ld_x16(col + 48, r1);
      tmem_wait_ld();
      split16(r0, out, skip);
      tmem_st_x16(col + 32, out);
      split16(r1, out, skip);
      tmem_st_x16(col + 48, out);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster_relaxed(bar0_lead + 8 * (kAFull + slot)); else mbar_arrive(bar0 + 8 * (kAFull + slot)); }
      if (q == 0) trace_event(p, kTrProdArrived + 2 * group, n);
      slot += NPG;
      if (slot >= ring) { slot -= ring; ph ^= 1; }
    }
    }
  } else {
    // ------------------------------------------------------------------ candidate rows + epilogue
    // warps 0..3 (set 0) and, with PAIR, warps kMainWarp + 2 .. + 5 (set 1); a warp reaches the TMEM lanes of its id mod 4
    const int q = warp & 3;                          // <-> TMEM lanes 32q..32q+31
    const int eset = (warp >= 4) ? 1 : 0;            // set 0 also stages the candidate rows; the sets alternate column blocks
    const uint32_t lane_base = static_cast<uint32_t>(q * 32) << 16;

    // Write the split rows of this CTA's j-th item into buffer j % xa_bufs; called after the epilogue of item
    // j - xa_bufs, when every projection that read the buffer has long completed.  K order per input dimension i:
    // x1 x1 x2 x1 x2 x3 (against w1 w2 w1 w3 w2 w1), then three ones (against b1 b2 b3), zero padding.
    auto stage_rows = [&](int j) {
      if (unit + j * n_units >= n_items) return;
      const int item = item_of(unit + j * n_units);
      const int xb = (xa_bufs == 2) ? (j & 1) : 0;
      // PAIR: this CTA's tile of the pair; an odd tile count leaves the peer's last tile empty (clamped here, skipped in
      // the epilogue)
      const int mtile = PAIR ? min(2 * (item / n_chunks) + static_cast<int>(cta), p.n_mtiles - 1) : item / n_chunks;
      // row of X holding input dimensions [0, dsplit) and the row holding [dsplit, d): the same row for plain
      // candidates; x_i and x_j for duels (tile = 128 consecutive i of one j)
      long long row_a, row_b;
      int dsplit;
      if (p.duel_n > 0) {
        const int tpi = (p.duel_nj + kTileM - 1) / kTileM;
        row_a = mtile / tpi;
        row_b = p.duel_j0 + min((mtile % tpi) * kTileM + q * 32 + lane, p.duel_nj - 1);
        dsplit = p.d >> 1;
      } else {
        row_a = min(static_cast<long long>(mtile) * kTileM + q * 32 + lane, p.N - 1);   // tail rows are masked in the epilogue
        row_b = row_a;
        dsplit = p.d;
      }
      if (p.ready_flag != nullptr) {
        // tiles never straddle chunks (rows_per_chunk is a multiple of 128); the copy engine owes us chunk `need - 1`
        const unsigned need = static_cast<unsigned>(row_a / p.rows_per_chunk) + 1u;
        if (lane == 0) {
          const long long t0 = clock64();
          for (;;) {
            unsigned got;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(p.ready_flag) : "memory");
            if (got >= need) break;
            // a wait that has expired anywhere (this or another warp / CTA) ends every later wait at once: the launch is
            // already lost, the host reports it, so the total stall is bounded by ONE time limit (~10 s), not one per tile
            unsigned lost;
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(lost) : "l"(p.ready_flag + 1) : "memory");
            if (lost != 0u) break;
            if (clock64() - t0 > 20000000000ll) { atomicExch(p.ready_flag + 1, 1u); break; }
            __nanosleep(100);
          }
        }
        __syncwarp();
      }
      const int dx = (p.duel_n > 0) ? dsplit : p.d;  // row stride of X
      uint32_t w[kPCols];
#pragma unroll
      for (int c = 0; c < kPCols; ++c) w[c] = 0u;
#pragma unroll
      for (int i = 0; i < DT; ++i) {
        // L2 loads: every element is read once, and streamed rows must not be looked up in a non-coherent cache
        const float x = (i < p.d) ? ((i < dsplit) ? __ldcg(p.X + row_a * dx + i) : __ldcg(p.X + row_b * dx + (i - dsplit))) : 0.f;
        uint32_t b1, b2, b3;
        bf16_split3(x, b1, b2, b3);
        w[3 * i + 0] = b1 | (b1 << 16);              // k = 6i, 6i+1
        w[3 * i + 1] = b2 | (b1 << 16);              // k = 6i+2, 6i+3
        w[3 * i + 2] = b2 | (b3 << 16);              // k = 6i+4, 6i+5
      }
      constexpr uint32_t kOne = 0x3F80u;             // bf16 1.0 at k = 6 DT, 6 DT + 1, 6 DT + 2
      w[3 * DT] = kOne | (kOne << 16);
      w[3 * DT + 1] = kOne;
      if constexpr (kXSmem) {
        // row r of a K-major tile with kPRow-byte rows and the matching swizzle (what TMA writes for the (W | b)
        // tile): 16-byte chunk c of the row lands at chunk c ^ f(r), f = r & 7 / (r >> 1) & 3 / (r >> 2) & 1 for
        // 128 / 64 / 32-byte rows
        const int r = q * 32 + lane;
        const uint32_t sw = (kPRow == 128) ? (r & 7) : ((kPRow == 64) ? ((r >> 1) & 3) : ((r >> 2) & 1));
        const uint32_t rowaddr = xa_smem0 + xb * kXBytes + r * kPRow;
#pragma unroll
        for (int c = 0; c < kPCols / 4; ++c) {
          constexpr int kChunksPerRow = kPRow / 16;
          const uint32_t tile = c / kChunksPerRow, cc = c % kChunksPerRow;
          const uint32_t addr = rowaddr + tile * (kTileM * kPRow) + ((cc ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[4 * c]), "r"(w[4 * c + 1]),
                       "r"(w[4 * c + 2]), "r"(w[4 * c + 3]) : "memory");
        }
        fence_proxy_async_smem();
      } else {
        const uint32_t col = tmem_base + lane_base + xa_col0 + xb * kPCols;
#pragma unroll
        for (int c = 0; c < kPCols; c += 8) tmem_st_x8(col + c, w + c);
        tmem_wait_st();
        tc_fence_before();
      }
      __syncwarp();
      if (lane == 0) { if constexpr (PAIR) mbar_arrive_cluster(bar0_lead + 8 * (kXaFull + xb)); else mbar_arrive(bar0 + 8 * (kXaFull + xb)); }
    };

    if (eset == 0) {
      stage_rows(0);
      if (xa_bufs == 2) stage_rows(1);
    }

    int ab = 0, nitem = 0;
    uint32_t aph = 0;
    // MODE 0: thr_s[c] is a lower bound of the running maximum of sample s_base + c over everything any CTA has recorded
    // so far (the global keys, re-read at every item, plus this CTA's own finds).  A block of 32 columns in which no row
    // reaches its column's bound cannot change the result and costs one compare per value; the full (max, lowest row)
    // reduction + atomicMax runs only for blocks that may improve a sample -- after the first tiles that is rare
    // (a sample's maximum improves ~ln(#tiles) times).  Ties go through the slow path: the key's ~row decides.

    for (int it_i = unit; it_i < n_items; it_i += n_units, ++nitem) {
      const int item = item_of(it_i);
      const int mtile = PAIR ? 2 * (item / n_chunks) + static_cast<int>(cta) : item / n_chunks;
      const int chunk = item % n_chunks;
      const int s_base = chunk * SC;
      // per-column constants of this item's chunk in shared memory (thr_s): the running-maximum bounds (MODE 0) or the
      // per-sample unscale factors (MODE 1, 2); loads issued before the wait, stored after it
      float thr_new[2];
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = (eset == 0) ? t * 128 + q * 32 + lane : SC;      // set 0 loads and publishes the per-column constants
        if constexpr (MODE == 0) {
          const unsigned long long pk = (c < SC) ? __ldcg(p.packed + s_base + c) : 0ull;
          thr_new[t] = (pk == 0ull) ? -INFINITY : f32_from_orderable(static_cast<uint32_t>(pk >> 32));
        } else {
          thr_new[t] = (c < SC) ? __ldg(p.unscale + s_base + c) : 0.f;
        }
      }
      mbar_wait(bar0 + 8 * (kAccFull + ab), aph);
      // the epilogue warps share thr_s: nobody still reads the previous item's values / everybody sees the new ones
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiSets) : "memory");
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int c = (eset == 0) ? t * 128 + q * 32 + lane : SC;
        if (c < SC) thr_s[c] = thr_new[t];
      }
      asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiSets) : "memory");
      if (warp == 0) trace_event(p, kTrEpiAccFull, nitem);
      tc_fence_after();
      const long long row0 = static_cast<long long>(mtile) * kTileM + q * 32;
      const long long row = row0 + lane;
      const bool partial = (row0 + 32 > p.N);
      const bool row_ok = row < p.N;
      const uint32_t acc = tmem_base + lane_base + ab * SC;
      const bool halves = PAIR && p.halves;
      const int half_blocks = SC / 64;                 // 32-column blocks per accumulator half (halves mode)
      // halves mode: blocks [0, half_blocks) first, then acc_empty[0] is released while the second half is read
#pragma unroll 1
      for (int part = 0; part < (halves ? 2 : 1); ++part) {
      if (part == 1) {                                 // half 1 has its own completion (it may trail half 0)
        mbar_wait(bar0 + 8 * (kAccFull + 1), aph);
        tc_fence_after();
      }
      const int j_begin = halves ? part * half_blocks + eset : eset;
      const int j_end = halves ? (part + 1) * half_blocks : SC / 32;
      // this warp's share of the accumulator (half) is released as soon as its last block sits in registers, i.e. before
      // that block's arithmetic (sigmoids + reductions in duel mode, the stores of store-F)
      bool released = false;
      auto release_acc = [&]() {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (PAIR) mbar_arrive_cluster_relaxed(bar0_lead + 8 * (kAccEmpty + (halves ? part : ab)));
          else mbar_arrive(bar0 + 8 * (kAccEmpty + ab));
        }
        released = true;
      };
      if (row0 < p.N && mtile < p.n_mtiles) {        // else: whole warp beyond the last candidate (warp-uniform)
#pragma unroll 1
        for (int j = j_begin; j < j_end; j += kEpiSets) {
          uint32_t v[32];
          tmem_ld_x32(acc + j * 32, v);
          tmem_wait_ld();
          if (j + kEpiSets >= j_end) release_acc();
          if constexpr (MODE == 0) {
            if (partial) {
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = row_ok ? v[c] : 0xFF800000u;     // -inf
            }
            if (approx_pass) {
              // approximate maximum of this warp's 32 rows in each of the block's 32 columns (NaN rows drop out of the
              // max; an all-NaN group yields NaN, which the marking kernel treats as "cannot be excluded")
              float gm = -INFINITY;
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const float m = warp_max_f32(__uint_as_float(v[c]));
                if (lane == c) gm = m;
              }
              p.rg_max[(static_cast<long long>(mtile) * 4 + q) * (static_cast<long long>(SC) * n_chunks) + s_base + j * 32 + lane] = gm;
            }
            bool reach = false;                        // !(val < bound): also true for NaN, which the slow path skips
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 t4 = *reinterpret_cast<const float4*>(&thr_s[j * 32 + c4 * 4]);
              reach |= !(__uint_as_float(v[4 * c4 + 0]) < t4.x);
              reach |= !(__uint_as_float(v[4 * c4 + 1]) < t4.y);
              reach |= !(__uint_as_float(v[4 * c4 + 2]) < t4.z);
              reach |= !(__uint_as_float(v[4 * c4 + 3]) < t4.w);
            }
            if (__any_sync(0xffffffffu, reach)) {
              float mine = -INFINITY;
              unsigned minehit = 0u;
#pragma unroll
              for (int c = 0; c < 32; ++c) {
                const float val = __uint_as_float(v[c]);
                const float m = warp_max_f32(val);
                const unsigned hit = __ballot_sync(0xffffffffu, val == m);
                if (lane == c) { mine = m; minehit = hit; }
              }
              // lane c holds (max, rows attaining it) of column c; record it if it may beat (or tie) the bound
              const int cs = j * 32 + lane;
              if (minehit != 0u && mine > -INFINITY && !(mine < thr_s[cs]) && s_base + cs < p.S) {
                if (approx_pass) {
                  // pass 1: the recorded bound is 2 eps below the largest approximate value: a value that reaches it may,
                  // in full precision, reach the largest value seen
                  const float lb = mine - __ldg(p.eps2 + s_base + cs);
                  atomicMax(p.packed + s_base + cs, pack_val_row(lb + 0.0f, static_cast<uint32_t>(row0) + (__ffs(minehit) - 1)));
                  if (!(lb < thr_s[cs])) thr_s[cs] = lb;
                } else {
                  atomicMax(p.packed + s_base + cs,
                            pack_val_row(mine + 0.0f, static_cast<uint32_t>(row0) + (__ffs(minehit) - 1)));
                  thr_s[cs] = mine;                    // racing warps may leave the lower of two finds: still a lower bound
                }
              }
              __syncwarp();
            }
          } else if constexpr (MODE == 1) {
            if (row_ok) {
              // column c of the block -> row s0 + c of out_F: 32 lanes write 128 contiguous bytes.  Whole blocks (the usual
              // case: S is a multiple of 32) take a branch-free path - unscale factors as float4, one pointer bumped by ldf -
              // so that the 32 multiplies and stores pipeline; with a test per column the read-out took ~3400 cycles per
              // block and the store-F kernel spent a third of its time waiting for the accumulator (ncu: tensor pipe 66 %)
              const int s0 = s_base + j * 32;
              float* dst = p.out_F + static_cast<long long>(s0) * p.ldf + row;
              if (s0 + 32 <= p.S) {
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                  const float4 t4 = *reinterpret_cast<const float4*>(&thr_s[j * 32 + c4 * 4]);
                  dst[0] = __uint_as_float(v[4 * c4 + 0]) * t4.x;
                  dst[p.ldf] = __uint_as_float(v[4 * c4 + 1]) * t4.y;
                  dst[2 * p.ldf] = __uint_as_float(v[4 * c4 + 2]) * t4.z;
                  dst[3 * p.ldf] = __uint_as_float(v[4 * c4 + 3]) * t4.w;
                  dst += 4 * p.ldf;
                }
              } else {
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                  if (s0 + c < p.S) dst[static_cast<long long>(c) * p.ldf] = __uint_as_float(v[c]) * thr_s[j * 32 + c];
                }
              }
            }
          } else {
            // soft-Copeland partial sums (dts.py:94-96): this tile holds f_s([x_i, x_j]) for one i and 128 values of j.
            // sigmoid in fp32, rounded to 2^-23 fixed point (kCopelandFixBits), summed over the warp's 32 rows with an integer
            // redux and added to score_fix[s, i] with one 64-bit atomic per (warp, sample): integer sums commute exactly.
            // The read-out is not overlapped when the accumulator is single-buffered and its XU work (ex2, rcp, float ->
            // int) bound it, so: two columns share one reciprocal, 1/((1+e0)(1+e1)) times the other factor (f clamped
            // to +-20, beyond which the sigmoid rounds to 0 or 1 on the grid anyway, keeps the product finite), and the
            // rounding to integer is the magic-number add  s 2^23 + 2^23  on the FMA pipe (exact for 0 <= s <= 1).
            const int tpi = (p.duel_nj + kTileM - 1) / kTileM;
            const int i_row = mtile / tpi;
            const bool j_ok = (mtile % tpi) * kTileM + q * 32 + lane < p.duel_nj;
            // two passes: all 32 fixed-point values first (independent arithmetic the scheduler can interleave), then the 32
            // warp reductions -- a redux.sync between the columns would serialise their ex2 -> rcp -> fma chains
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              float d[2];
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const float f = fminf(fmaxf(__uint_as_float(v[c + t]) * thr_s[j * 32 + c + t], -20.f), 20.f);
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * f));
                d[t] = 1.f + e;
              }
              float rinv;
              asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(d[0] * d[1]));
#pragma unroll
              for (int t = 0; t < 2; ++t) {
                const float sg = rinv * d[1 - t];                                     // 1 / d[t]
                const unsigned bits = __float_as_uint(fmaf(sg, kCopelandFixScale, kCopelandFixScale));
                v[c + t] = j_ok ? bits - 0x4B000000u : 0u;                            // round(sg 2^23) = mantissa of sg 2^23 + 2^23
              }
            }
            unsigned mine = 0u;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
              const unsigned r = __reduce_add_sync(0xffffffffu, v[c]);
              if (lane == c) mine = r;
            }
            const int s = s_base + j * 32 + lane;
            if (s < p.S)
              atomicAdd(p.score_fix + static_cast<long long>(s) * p.duel_n + i_row, static_cast<unsigned long long>(mine));
          }
        }
      }
      if (!released) release_acc();
      }   // part
      if (warp == 0) trace_event(p, kTrEpiDone, nitem);
      if (++ab == acc_bufs) { ab = 0; aph ^= 1; }
      if (eset == 0) stage_rows(nitem + xa_bufs);
    }
  }

  // ---------------------------------------------------------------------- teardown + final reduce
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();    // nobody signals the peer's barriers / TMEM any more
  tc_fence_after();
  if (warp == kMainWarp) {
    if constexpr (PAIR) tmem_dealloc_pair(tmem_base, kTmemCols); else tmem_dealloc(tmem_base, kTmemCols);
  }
  if constexpr (MODE == 0) {
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
      const unsigned prev = atomicAdd(p.done_counter, 1u);
      is_last_cta = (prev == gridDim.x - 1);
      peer_timed_out = 0;
    }
    __syncthreads();
    if (is_last_cta && approx_pass) {
      // pass 1: nothing to report yet.  Hand the final bounds to the marking kernel and clear the keys (pass 2 starts from
      // scratch with exact values)
      __threadfence();
      for (int s = threadIdx.x; s < p.SC * n_chunks; s += blockDim.x) {
        const unsigned long long pk = __ldcg(p.packed + s);
        p.lb_out[s] = (pk == 0ull) ? -INFINITY : f32_from_orderable(static_cast<uint32_t>(pk >> 32));
        p.packed[s] = 0ull;
      }
      if (threadIdx.x == 0) *p.done_counter = 0u;
    } else if (is_last_cta) {
      __threadfence();
      for (int s = threadIdx.x; s < p.S; s += blockDim.x) {
        const unsigned long long pk = __ldcg(p.packed + s);
        p.packed[s] = 0ull;                       // ready for the next launch
        const bool none = (pk == 0ull);           // only NaNs seen for this sample
        const float v = none ? __int_as_float(0x7fc00000)
                             : f32_from_orderable(static_cast<uint32_t>(pk >> 32)) * p.unscale[s];
        const long long gi = none ? -1 : static_cast<long long>(0xFFFFFFFFu - static_cast<uint32_t>(pk)) + p.idx_offset;
        if (p.out_key != nullptr || p.peer_world > 0) {
          const long long key = none ? kEmptyKey : argmax_key(v, static_cast<uint32_t>(gi));
          if (p.out_key != nullptr) p.out_key[s] = key;
          const size_t slot = (static_cast<size_t>(p.peer_epoch & 1ull) * p.peer_world + p.peer_rank) * p.S + s;
          // remote ranks first: their merges wait for us over NVLink, ours reads the local copy last
          for (int r = 1; r <= p.peer_world; ++r) {
            const int dst = (p.peer_rank + r) % p.peer_world;
            p.peer_buf[dst][slot] = static_cast<unsigned long long>(key);
          }
        } else {
          p.out_val[s] = v;
          p.out_idx[s] = gi;
        }
      }
      if (threadIdx.x == 0) *p.done_counter = 0u;
      if (p.peer_world > 0) {
        __threadfence_system();                     // every thread's key stores before the flags
        __syncthreads();
        if (threadIdx.x < p.peer_world) {
          unsigned long long* flag = p.peer_buf[threadIdx.x] + 2ull * p.peer_world * p.S + p.peer_rank;
          asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(p.peer_epoch) : "memory");
        }
        if (p.peer_merge) {
          // merge here instead of in a second launch: acquire every rank's flag in the local buffer (bounded wait), then
          // per sample the maximum over the `world` slots; max commutes, so every rank computes the same bits
          const unsigned long long* local = p.peer_buf[p.peer_rank];
          if (threadIdx.x < p.peer_world) {
            const unsigned long long* flag = local + 2ull * p.peer_world * p.S + threadIdx.x;
            const long long t0 = clock64();
            for (;;) {
              unsigned long long f;
              asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(flag) : "memory");
              if (f >= p.peer_epoch) break;
              if (clock64() - t0 > 20000000000ll) { peer_timed_out = 1; break; }
              __nanosleep(64);
            }
          }
          __syncthreads();
          const unsigned long long* keys = local + static_cast<size_t>(p.peer_epoch & 1ull) * p.peer_world * p.S;
          for (int s = threadIdx.x; s < p.S; s += blockDim.x) {
            if (peer_timed_out) { p.out_val[s] = __int_as_float(0x7fc00000); p.out_idx[s] = -2; continue; }
            long long best = kEmptyKey;
            for (int r = 0; r < p.peer_world; ++r) {
              unsigned long long k;
              asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(k) : "l"(keys + static_cast<size_t>(r) * p.S + s) : "memory");
              best = max(best, static_cast<long long>(k));
            }
            float bv;
            long long bi;
            argmax_unkey(best, &bv, &bi);
            p.out_val[s] = bv;
            p.out_idx[s] = bi;
          }
          if (peer_timed_out && threadIdx.x == 0 && p.peer_status != nullptr)
            *reinterpret_cast<volatile unsigned int*>(p.peer_status) = static_cast<unsigned int>(p.peer_epoch);
        }
      }
    }
  }
}

}  // namespace tkbo
