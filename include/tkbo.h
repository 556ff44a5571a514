/*
 * tkbo.h — C ABI of libtkbo.so: B200 (sm_100a) kernels for the one data-parallel hot path of
 * sebtsh/Top-k-Ranking-Bayesian-Optimization: random-Fourier-feature (RFF) posterior sampling with a
 * fused per-sample argmax, and the MPES Monte-Carlo ranking-likelihood entropy estimate.
 *
 * The reference is pure Python and has no FFI of its own; its boundary is the Python call signatures
 * in fourier_features.py and acquisitions/{pes,rank_pes,dts}.py.  Each entry point below names the
 * reference lines whose arithmetic it replaces; INTEGRATION.md shows the ctypes stub a maintainer of
 * the reference would add.
 *
 * Conventions
 *   - plain C symbols, no C++/torch types; every function returns 0 on success or a negative
 *     tkbo_status code, never throws; tkbo_last_error() returns a thread-local message.
 *   - unless a function name ends in _host, every data pointer is a DEVICE pointer owned by the caller
 *     (row-major, densely packed unless a leading dimension is given); calls are asynchronous on
 *     `stream` (a cudaStream_t passed as void*; NULL = default stream).
 *   - one handle per device; a handle is not thread-safe (the reference is single-threaded).
 *   - there is no CPU fallback: on a machine without an sm_100 GPU tkbo_create fails.
 */
#ifndef TKBO_H_
#define TKBO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TKBO_VERSION 100  /* 0.1.0 */

typedef enum {
  TKBO_OK = 0,
  TKBO_ERR_INVALID = -1,      /* bad argument (shape, null pointer, unsupported size) */
  TKBO_ERR_CUDA = -2,         /* CUDA runtime / driver error, message in tkbo_last_error() */
  TKBO_ERR_UNSUPPORTED = -3,  /* device is not sm_100, or configuration outside the kernel's limits */
  TKBO_ERR_NOMEM = -4
} tkbo_status;

typedef struct tkbo_handle_s* tkbo_handle_t;
typedef struct tkbo_basis_s* tkbo_basis_t;

/* ---- lifecycle ----------------------------------------------------------------------------------- */
int tkbo_version(void);
const char* tkbo_last_error(void);
/* Binds to CUDA device `device`; fails with TKBO_ERR_UNSUPPORTED unless it is compute capability 10.x. */
int tkbo_create(tkbo_handle_t* out, int device);
int tkbo_destroy(tkbo_handle_t h);
/* Number of SMs the persistent kernels are sized for (148 on B200). */
int tkbo_sm_count(tkbo_handle_t h);

/* ---- shared-basis RFF sampler (kernel K1) ---------------------------------------------------------
 * One basis (W, b) shared by S posterior weight samples Theta (D, S):
 *     F[n, s] = sqrt(2/D) * sum_j cos(X[n,:].W[j,:] + b[j]) * Theta[j, s]
 * which is fourier_features(X, W, b) @ theta  (fourier_features.py:18-21 and :162; acquisitions/dts.py:82-85)
 * evaluated for all S samples at once.  The basis object holds the device-side operand layout
 * (bf16-split (W | b) projection rows, Theta split into fp16 hi/lo K-major tiles with a per-sample power-of-two scale). */
/* Calls that use the same basis object must be stream-ordered (it owns the argmax workspace); calls on different basis
 * objects may run concurrently on different streams. */
int tkbo_rff_basis_create(tkbo_handle_t h, tkbo_basis_t* out, int D, int d, int S);
/* Same with an explicit contraction precision.  Phi and Theta are each split into an fp16 hi part and a remainder:
 *   TKBO_PRECISION_FP32  hi*hi + hi*lo + lo*hi as three fp16 tcgen05.mma (fp32 emulation, ~2e-6 of max|F|);
 *   TKBO_PRECISION_FAST  hi*hi in fp16 + both correction terms in ONE e4m3 tcgen05.mma (K = 32): two thirds of the
 *                        tensor work; error ~1e-5 ||sqrt(2/D) theta_s||_2, i.e. ~3e-5 of max|F| for a well-posed
 *                        regression and above the path's 1e-4 tolerance when the weights cancel (||theta'|| >> max|F|);
 *   TKBO_PRECISION_AUTO  (what tkbo_rff_basis_create and the _host entry points use) = FP32: the library does not know
 *                        max|F|, so it never trades accuracy on its own.  The Python SharedBasis picks FAST only when the
 *                        caller supplies the scale of f and the bound stays below 0.3e-4 of it. */
#define TKBO_PRECISION_AUTO 0
#define TKBO_PRECISION_FP32 1
#define TKBO_PRECISION_FAST 2
int tkbo_rff_basis_create_ex(tkbo_handle_t h, tkbo_basis_t* out, int D, int d, int S, int precision);
int tkbo_rff_basis_destroy(tkbo_handle_t h, tkbo_basis_t basis);
/* W [D, d], b [D], Theta [D, S] (theta_is_transposed = 0) or [S, D] (= 1); all float64 device arrays. */
int tkbo_rff_basis_set(tkbo_handle_t h, tkbo_basis_t basis, const double* W, const double* b,
                       const double* Theta, int theta_is_transposed, void* stream);

/* out_val[s] = max_n F[n, s], out_idx[s] = idx_offset + (lowest n attaining it)  — the argmax + gather
 * of fourier_features.py:162-167 with n_init -> N, without ever writing F.  X [N, d] float32.
 * N < 2^32.  idx_offset lets a caller shard the candidate axis across GPUs. */
int tkbo_rff_argmax(tkbo_handle_t h, tkbo_basis_t basis, const float* X, int64_t N, int64_t idx_offset,
                    float* out_val, int64_t* out_idx, void* stream);
/* out_F[s * ldf + n] = F[n, s] (sample-major, ldf >= N): the (N, count) matrix of
 * acquisitions/dts.py:85 / the f-samples consumed by tkbo_mpes_topk. */
int tkbo_rff_eval(tkbo_handle_t h, tkbo_basis_t basis, const float* X, int64_t N, float* out_F,
                  int64_t ldf, void* stream);
/* Dueling-Thompson soft-Copeland scores without materialising the n^2 duels (acquisitions/dts.py:31-45 combinations,
 * :71-85 sample_f, :88-97 soft_copeland_maximizer): the basis has d = 2 d_obj, the candidates are all ordered pairs
 * [x_i, x_j] of the n base points Xbase [n, d_obj]; out_score[s, i] = mean_j sigmoid(f_s([x_i, x_j])) (float32 [S, n],
 * may be NULL) and out_argmax[s] = lowest i attaining the maximum (may be NULL).  The per-duel values never leave the SM. */
int tkbo_rff_duel_copeland(tkbo_handle_t h, tkbo_basis_t basis, const float* Xbase, int n, float* out_score,
                           int64_t* out_argmax, void* stream);
/* Sharded form: accumulate fix[s, i] += sum_{j in [j_begin, j_end)} round(2^23 sigmoid(f_s([x_i, x_j]))) into the caller's
 * zero-initialised uint64 [S, n] buffer (integer sums: shards held by different ranks combine exactly with one
 * all-reduce(SUM) over int64), then tkbo_copeland_finish turns the sums into mean scores and winners. */
int tkbo_rff_duel_copeland_partial(tkbo_handle_t h, tkbo_basis_t basis, const float* Xbase, int n, int j_begin, int j_end,
                                   uint64_t* fix, void* stream);
int tkbo_copeland_finish(tkbo_handle_t h, const uint64_t* fix, int S, int n, float* out_score, int64_t* out_argmax,
                         void* stream);
/* Same as tkbo_rff_argmax but returns one signed 64-bit key per sample whose integer maximum is "largest value, then
 * lowest global index" (idx_offset + N < 2^32): ranks holding candidate shards merge with ONE all-reduce(MAX) over
 * int64 (NCCL), then tkbo_unpack_argmax_key recovers (value, index).  An empty / all-NaN sample has key INT64_MIN. */
int tkbo_rff_argmax_key(tkbo_handle_t h, tkbo_basis_t basis, const float* X, int64_t N, int64_t idx_offset,
                        int64_t* out_key, void* stream);
int tkbo_unpack_argmax_key(tkbo_handle_t h, const int64_t* key, int S, float* out_val, int64_t* out_idx, void* stream);
/* The same merge fused into the kernel over NVLink peer memory (no collective call; SURVEY 8(e)'s all-gather of 8 S bytes
 * per rank done by the producing kernel): every rank owns a peer buffer of tkbo_peer_buffer_bytes(world, S) bytes, zeroed
 * once, that ALL ranks of the node can store to (peer_bufs[r] = rank r's buffer as mapped in this process: CUDA IPC /
 * torch symmetric memory).  tkbo_rff_argmax_peer runs the fused kernel and has its last CTA store the S keys into slot
 * (epoch & 1, rank) of every rank's buffer, then release flag[rank] = epoch there; tkbo_merge_peer_keys (same stream,
 * same epoch, on every rank) acquires all `world` flags of the LOCAL buffer, takes the per-sample maximum and unpacks.
 * Epochs start at 1 and increase by one per call pair on all ranks alike; world <= 8 (one NVSwitch node).  If a peer
 * never delivers, the merge gives up after ~10 s and returns index -2 for every sample instead of hanging. */
int64_t tkbo_peer_buffer_bytes(int world, int S);
/* out_val / out_idx both non-NULL: the launch's last CTA also performs the merge (it acquires the `world` flags of this
 * rank's own buffer after delivering its keys) and writes the GLOBAL (value, index) there -- one launch per rank and
 * iteration, no tkbo_merge_peer_keys call.  Both NULL: delivery only, merge with tkbo_merge_peer_keys.
 * All calls of one peer group must be issued in the same order on every rank (epochs advance in lock step) and, per rank,
 * on one stream. */
int tkbo_rff_argmax_peer(tkbo_handle_t h, tkbo_basis_t basis, const float* X, int64_t N, int64_t idx_offset,
                         void* const* peer_bufs, int world, int rank, uint64_t epoch, float* out_val, int64_t* out_idx,
                         void* stream);
/* 0, or the epoch of a merge on this handle that gave up waiting for a peer (index -2 in its outputs); reading clears it.
 * The word lives in host-mapped memory, so it is meaningful once the stream that ran the merge has been synchronised. */
int64_t tkbo_peer_status(tkbo_handle_t h);
/* Diagnostics of the two-pass argmax (every tkbo_rff_argmax* entry point, large problems): after the stream of the last call
 * has been synchronised, *listed = work items the full-precision pass covered, *total = work items in all; both -1 if that
 * call ran as a single pass.  The results do not depend on the pass structure. */
int tkbo_two_pass_items(tkbo_handle_t h, int* listed, int* total);
/* Delivers ready-made keys [S] for this epoch instead of running the fused kernel; keys == NULL delivers "no candidate"
 * (INT64_MIN) -- what a rank whose shard is empty calls so that its peers' merges do not wait for it. */
int tkbo_peer_push_keys(tkbo_handle_t h, const int64_t* keys, int S, void* const* peer_bufs, int world, int rank,
                        uint64_t epoch, void* stream);
int tkbo_merge_peer_keys(tkbo_handle_t h, const void* local_buf, int world, int S, uint64_t epoch, float* out_val,
                         int64_t* out_idx, void* stream);
/* Per-rank partial results [G, S] (e.g. after an all-gather) -> global (value, lowest index). */
int tkbo_merge_argmax(tkbo_handle_t h, const float* vals, const int64_t* idx, int G, int S,
                      float* out_val, int64_t* out_idx, void* stream);

/* Host-buffer convenience wrapper around basis_set + argmax (copies inside, synchronous):
 * the end-to-end call a non-torch caller would make. All arrays are HOST pointers.  `precision` = TKBO_PRECISION_* as for
 * tkbo_rff_basis_create_ex (AUTO = FP32; a caller that can bound ||theta'|| / max|f| passes FAST).  The fused kernel is launched once,
 * right after the basis is prepared, and consumes the candidates in up to 16 chunks while their H2D copies (second
 * stream) are still landing -- pass pinned X for that overlap; pageable X works, its copies then run host-synchronously
 * behind the launch.  Returns TKBO_ERR_CUDA if a chunk fails to arrive within ~10 s (the kernel does not hang).
 * Profilers / sanitizers that serialise kernels against copies defeat the overlap (every launch would sit out its wait
 * limit): set TKBO_HOST_STREAMED=0 in the environment to queue all copies before the launch. */
int tkbo_rff_argmax_host(tkbo_handle_t h, const double* W, const double* b, const double* Theta,
                         int D, int d, int S, const float* X, int64_t N, int precision, float* out_val, int64_t* out_idx);
/* The same for one rank of a candidate-sharded job on one NVSwitch node: X [N, d] is this rank's shard (rows idx_offset ..
 * idx_offset + N of the global set), peer_bufs / world / rank / epoch as in tkbo_rff_argmax_peer.  One fused launch per call:
 * it streams the shard in, delivers this rank's keys to every peer, merges in its last CTA, and the GLOBAL S results come
 * back with a single D2H copy -- no collective call, no second launch.  Synchronous; returns TKBO_ERR_CUDA if a chunk or a
 * peer fails to arrive within ~10 s.  Every rank of the group must make the call with the same epoch. */
int tkbo_rff_argmax_host_peer(tkbo_handle_t h, const double* W, const double* b, const double* Theta, int D, int d, int S,
                              const float* X, int64_t N, int precision, int64_t idx_offset, void* const* peer_bufs, int world,
                              int rank, uint64_t epoch, float* out_val, int64_t* out_idx);

/* ---- per-sample-basis RFF (reference-exact shapes, float64 end to end) ----------------------------
 * X [count, n, d], W [count, D, d], b [count, D], theta [count, D]:
 *     out_f[c, i] = sqrt(2/D) * sum_j cos(X[c,i,:].W[c,j,:] + b[c,j]) * theta[c, j]     (fourier_features.py:162)
 *     out_idx[c]  = argmax_i out_f[c, i], lowest index on ties                          (fourier_features.py:164-165)
 * out_f may be NULL. */
int tkbo_rff_eval_batched(tkbo_handle_t h, const double* X, const double* W, const double* b,
                          const double* theta, int count, int n, int D, int d, double* out_f,
                          int64_t* out_idx, void* stream);
/* out_phi[c, i, j] = sqrt(2/D) cos(X[c,i,:].W[c,j,:] + b[c,j]) — the materialised feature map of
 * fourier_features.py:7-21 (used on the n training / inducing inputs, never on the candidate grid). */
int tkbo_rff_features_batched(tkbo_handle_t h, const double* X, const double* W, const double* b, int count,
                              int n, int D, int d, double* out_phi, void* stream);
/* fourier_features.py:133-157: sign-gradient ascent (Keras RMSprop rho=0: x <- clip(x - lr g/(|g|+eps)))
 * on loss = -mean_{c,i} f_c(x[c,i]); stops when |dloss/loss| < rel_tol or after num_steps.
 * X is updated in place; *steps_run (host int, may be NULL) receives the step count. Synchronous.
 * Runs in chunks of 64 steps (TKBO_ASCENT_CHUNK) without a grid-wide barrier: the update of a start point needs only its
 * own gradient, the global stop test is evaluated once per chunk from the recorded per-step losses, and a chunk that
 * contains the stop is repeated for the exact number of steps (deterministic, so the iterates are the same). */
int tkbo_rff_ascent_batched(tkbo_handle_t h, double* X, const double* W, const double* b, const double* theta,
                            int count, int n, int D, int d, double min_val, double max_val, int num_steps,
                            double lr, double eps, double rel_tol, int* steps_run, void* stream);

/* ---- MPES Monte-Carlo entropy estimate (kernel K2) -------------------------------------------------
 * fchi [S, Q, C] float32 f-samples at the query points, fstar_argmax [S] int32 in [0, M): index of the
 * maximiser that wins sample s (acquisitions/rank_pes.py:39), perms [P, k] int32 rows in ascending
 * preference (rank_pes.py:139).  out_mi [Q] float64.
 *   mode TKBO_MPES_RANK: acquisitions/rank_pes.py:58-105 (empty maximiser groups dropped).
 *   mode TKBO_MPES_PES : acquisitions/pes.py:66-123 (k must be 1, perms NULL; eps = 1e-7 added to the
 *                        group counts and likelihood sums, all M groups kept).
 * Calls on one handle share its scratch buffers and must be stream-ordered.
 * perms == NULL means "all k-permutations of C" (P ignored; rank_pes.py:55): register-resident fast path
 * for the common (C, k); an explicit table (P <= 1024 rows, e.g. the random 1000-row subset of
 * rank_pes.py:49-53) takes the table-driven kernel.  C <= 8, Q < 2^31. */
#define TKBO_MPES_RANK 0
#define TKBO_MPES_PES 1
int tkbo_mpes_topk(tkbo_handle_t h, const float* fchi, const int32_t* fstar_argmax, int S, int64_t Q,
                   int C, int M, const int32_t* perms, int P, int k, int mode, double* out_mi,
                   void* stream);
/* Top-1 MPES with an indifference outcome (acquisitions/indiff_pes.py:15-130): C + 1 observation types,
 * P(i) = e^{f_i} / (e^{f_i} + sum_{j != i} e^{f_j + threshold}), P(none) = 1 - sum_i P(i); same inputs / output as
 * tkbo_mpes_topk, maximisers that win no sample are dropped (indiff_pes.py:68-76).  C <= 8. */
int tkbo_mpes_indiff(tkbo_handle_t h, const float* fchi, const int32_t* fstar_argmax, int S, int64_t Q, int C, int M,
                     double threshold, double* out_mi, void* stream);
/* out[s] = argmax_m fstar[s, m] (lowest index), fstar [S, M] float32 — rank_pes.py:39 / pes.py:108. */
int tkbo_argmax_rows(tkbo_handle_t h, const float* fstar, int S, int M, int32_t* out, void* stream);

/* ---- DTS tail --------------------------------------------------------------------------------------
 * f [n*n] float32, i-major duels (acquisitions/dts.py:31-45): score[i] = mean_j sigmoid(f[i*n+j]),
 * *out_argmax = argmax_i score[i]   (acquisitions/dts.py:94-97).  out_score may be NULL. */
int tkbo_soft_copeland(tkbo_handle_t h, const float* f, int n, float* out_score, int64_t* out_argmax,
                       void* stream);

/* ---- introspection used by bench.py / tests --------------------------------------------------------
 * Number of kernel launches issued through this handle since creation. */
int64_t tkbo_launch_count(tkbo_handle_t h);
/* Describes the tiling chosen for a basis: fills cfg[0..15] = {S_chunk, n_chunks, n_slabs, smem stages,
 * accumulator buffers, d_template, smem_bytes, grid, TMEM ring slots, threads per CTA, candidate-row
 * buffers, producer warp groups, projection-issuing warps,
 * 1 if the e4m3 correction form is used, depth of the separate (W | b) tile ring or 0,
 * 1 if both producer groups convert every slab}; cfg must have room for 16 int32. */
int tkbo_rff_basis_config(tkbo_handle_t h, tkbo_basis_t basis, int64_t N, int32_t* cfg);
/* Diagnostics: while `trace` (device, uint64 [TKBO_TRACE_EVENTS][cap]) is set, CTA 0 of every K1 launch logs
 * clock64() at its pipeline events (rff_fused.cuh TraceEvent); pass NULL to switch it off. */
#define TKBO_TRACE_EVENTS 17
int tkbo_debug_set_trace(tkbo_handle_t h, uint64_t* trace, int cap);

#ifdef __cplusplus
}
#endif
#endif /* TKBO_H_ */
