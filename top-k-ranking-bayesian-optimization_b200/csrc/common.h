// Host-side state shared by the translation units of libtkbo.so.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/tkbo.h"

namespace tkbo {

void set_error(const std::string& msg);

#define TKBO_CUDA_CHECK(expr)                                                                      \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      ::tkbo::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                       \
      return TKBO_ERR_CUDA;                                                                        \
    }                                                                                              \
  } while (0)

#define TKBO_REQUIRE(cond, msg)                                                                    \
  do {                                                                                             \
    if (!(cond)) {                                                                                 \
      ::tkbo::set_error(std::string("invalid argument: ") + (msg));                                \
      return TKBO_ERR_INVALID;                                                                     \
    }                                                                                              \
  } while (0)

}  // namespace tkbo

struct tkbo_handle_s {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  int64_t launches = 0;
  unsigned int* done_counter = nullptr;   // device, zero between calls
  void* encode_tiled = nullptr;           // cuTensorMapEncodeTiled entry point
  // (K2 scratch is allocated per call from the stream-ordered pool; tkbo_create raises the pool's release threshold so
  // that steady-state calls reuse the same blocks without going back to the driver)
  // merge over peer memory: one host-mapped word that a merge (kernel tail or tkbo_merge_peer_keys) sets to the epoch it
  // gave up on; tkbo_peer_status() reads and clears it
  unsigned int* peer_status_host = nullptr;
  int* two_pass_host = nullptr;              // pinned [2]: items listed for pass 2 of the last two-pass argmax, items in all (-1: single pass)
  unsigned int* peer_status_dev = nullptr;
  // persistent state of the host-buffer entry point (tkbo_rff_argmax_host): basis + device staging, reused across calls
  struct HostPath {
    tkbo_basis_s* basis = nullptr;
    int D = 0, d = 0, S = 0, precision = 0;
    double* dbasis = nullptr;            // W | b | Theta as float64
    size_t dbasis_doubles = 0;
    float* dX = nullptr;                 // all N candidates; filled chunk by chunk while the kernel already runs
    size_t dX_floats = 0;
    // one device block  idx[S] i64 | val[S] f32 | flag[2] u32  read back with a single D2H copy into `hout` (pinned);
    // flag[0] = chunks landed so far, flag[1] = the kernel's "chunk never arrived" report
    unsigned char* dout = nullptr;
    unsigned char* hout = nullptr;
    unsigned int* hflag = nullptr;       // pinned: 1, 2, ..., kHostChunksMax (sources of the flag copies)
    int out_S = 0;
    cudaStream_t copy_stream = nullptr, exec_stream = nullptr;
    cudaEvent_t flag_reset = nullptr;
  } host;
  // optional K1 pipeline trace target (tkbo_debug_set_trace); caller-owned device memory
  unsigned long long* trace = nullptr;
  int trace_cap = 0;
};

struct tkbo_basis_s {
  int D = 0, d = 0, S = 0;
  int DT = 0;        // template input dimension (d padded up to a supported value)
  int Dpad = 0;      // multiple of 64
  int SC = 0;        // samples per chunk
  int n_chunks = 0;
  int Salloc = 0;    // n_chunks * SC
  int stages = 0;    // smem ring depth
  int ring = 0;      // TMEM ring depth
  int acc_bufs = 0;
  int xa_bufs = 0;
  int proj_warps = 1;
  int coop = 0;      // 1: the two producer groups share every slab (32 columns each)
  int wstages = 0;   // depth of the separate (W | b) tile ring (0: the tile travels in the Theta stage)
  int npg = 2;       // producer warp groups (2, 3 or 4)
  int pair = 0;      // 1: clusters of two CTAs share every tcgen05.mma (cta_group::2), each fetching half of the B operands
  int nsplit = 1;    // single CTAs: the two main MMA warps split the sample columns of a slab (0: one warp issues them all)
  int prune = 0;     // 1: MODE 0 launches run in two passes (RffParams::approx); TKBO_PRUNE
  int lead = 0;      // halves mode: slabs issued ahead per item (0 = ring depth); TKBO_LEAD
  int halves = 0;    // 1 (pairs, one accumulator buffer): the accumulator is retired in two column halves
  int corr8 = 0;     // 1: correction terms as one e4m3 MMA (fp16 hi*hi + fp8 [hi*lo | lo*hi]); 0: three fp16 MMAs
  int smem_bytes = 0;
  uint16_t* Wproj = nullptr;         // [Dpad, proj_k_alloc(DT)] bf16 split (W | b) rows, K order of rff_fused.cuh
  __half* theta_hi = nullptr;        // [Salloc, Dpad]
  __half* theta_lo = nullptr;        // [Salloc, Dpad] fp16 lo parts, or (corr8) [Salloc, 2 Dpad] e4m3 bytes: per 16
                                     // features 16 lo parts then 16 hi parts scaled by 2^-12
  float* unscale = nullptr;          // [Salloc]
  unsigned long long* packed = nullptr;  // [Salloc], zero between launches
  CUtensorMap tm_hi, tm_lo, tm_w;
  bool is_set = false;
};
