// extern "C" surface of libtkbo.so (declared in include/tkbo.h).  No torch, no exceptions across the ABI.
#include <cuda_fp16.h>

#include <new>
#include <string>

#include "common.h"

namespace tkbo {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

int basis_create(tkbo_handle_t h, tkbo_basis_t* out, int D, int d, int S);
int basis_destroy(tkbo_handle_t h, tkbo_basis_t bs);
int basis_set(tkbo_handle_t h, tkbo_basis_t bs, const double* W, const double* b, const double* Theta, int transposed,
              cudaStream_t stream);
int basis_config(tkbo_handle_t h, tkbo_basis_t bs, int64_t N, int32_t* cfg);
int rff_argmax(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, int64_t idx_offset, float* out_val,
               int64_t* out_idx, cudaStream_t stream);
int rff_eval(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, float* out_F, int64_t ldf, cudaStream_t stream);
int rff_argmax_key(tkbo_handle_t h, tkbo_basis_t bs, const float* X, int64_t N, int64_t idx_offset, int64_t* out_key,
                   cudaStream_t stream);
int unpack_argmax_key(tkbo_handle_t h, const int64_t* key, int S, float* out_val, int64_t* out_idx, cudaStream_t stream);
int merge_argmax(tkbo_handle_t h, const float* vals, const int64_t* idx, int G, int S, float* out_val, int64_t* out_idx,
                 cudaStream_t stream);
int rff_eval_batched(tkbo_handle_t h, const double* X, const double* W, const double* b, const double* theta, int count,
                     int n, int D, int d, double* out_f, int64_t* out_idx, cudaStream_t stream);
int rff_ascent_batched(tkbo_handle_t h, double* X, const double* W, const double* b, const double* theta, int count,
                       int n, int D, int d, double min_val, double max_val, int num_steps, double lr, double eps,
                       double rel_tol, int* steps_run, cudaStream_t stream);
int rff_features_batched(tkbo_handle_t h, const double* X, const double* W, const double* b, int count, int n, int D,
                         int d, double* out_phi, cudaStream_t stream);
int mpes_topk(tkbo_handle_t h, const float* fchi, const int32_t* fstar_argmax, int S, int64_t Q, int C, int M,
              const int32_t* perms, int P, int k, int mode, double* out_mi, cudaStream_t stream);
int argmax_rows(tkbo_handle_t h, const float* fstar, int S, int M, int32_t* out, cudaStream_t stream);
int soft_copeland(tkbo_handle_t h, const float* f, int n, float* out_score, int64_t* out_argmax, cudaStream_t stream);

}  // namespace tkbo

using namespace tkbo;

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Every entry point runs on the handle's device regardless of the caller's current device.
struct DeviceGuard {
  int prev = -1;
  bool ok = true;
  explicit DeviceGuard(tkbo_handle_t h) {
    if (!h) { ok = false; return; }
    if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
    if (prev != h->device && cudaSetDevice(h->device) != cudaSuccess) ok = false;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

#define TKBO_ENTER(h)                                              \
  if (!(h)) { set_error("null handle"); return TKBO_ERR_INVALID; } \
  DeviceGuard _guard(h);                                           \
  if (!_guard.ok) { set_error("cannot select the handle's device"); return TKBO_ERR_CUDA; }

extern "C" {

int tkbo_version(void) { return TKBO_VERSION; }

const char* tkbo_last_error(void) { return g_last_error.c_str(); }

int tkbo_create(tkbo_handle_t* out, int device) {
  if (!out) { set_error("null out pointer"); return TKBO_ERR_INVALID; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_error(std::string("no CUDA device available: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
              " (libtkbo has no CPU fallback)");
    return TKBO_ERR_UNSUPPORTED;
  }
  if (device < 0 || device >= ndev) { set_error("device index out of range"); return TKBO_ERR_INVALID; }
  cudaDeviceProp prop;
  TKBO_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    set_error(std::string("device '") + prop.name + "' is compute capability " + std::to_string(prop.major) + "." +
              std::to_string(prop.minor) + "; libtkbo is built for sm_100a only");
    return TKBO_ERR_UNSUPPORTED;
  }
  int prev = 0;
  TKBO_CUDA_CHECK(cudaGetDevice(&prev));
  TKBO_CUDA_CHECK(cudaSetDevice(device));
  auto* h = new (std::nothrow) tkbo_handle_s();
  if (!h) { set_error("out of host memory"); return TKBO_ERR_NOMEM; }
  h->device = device;
  h->sm_count = prop.multiProcessorCount;
  h->max_smem_optin = static_cast<int>(prop.sharedMemPerBlockOptin);
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &h->encode_tiled, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !h->encode_tiled) {
    delete h;
    cudaSetDevice(prev);
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return TKBO_ERR_UNSUPPORTED;
  }
  e = cudaMalloc(&h->done_counter, sizeof(unsigned int));
  if (e == cudaSuccess) e = cudaMemset(h->done_counter, 0, sizeof(unsigned int));
  if (e != cudaSuccess) {
    delete h;
    cudaSetDevice(prev);
    set_error(std::string("cudaMalloc (handle): ") + cudaGetErrorString(e));
    return TKBO_ERR_NOMEM;
  }
  cudaSetDevice(prev);
  *out = h;
  return TKBO_OK;
}

int tkbo_destroy(tkbo_handle_t h) {
  if (!h) return TKBO_OK;
  {
    DeviceGuard g(h);
    cudaFree(h->done_counter);
    cudaFree(h->mpes_scratch);
  }
  delete h;
  return TKBO_OK;
}

int tkbo_sm_count(tkbo_handle_t h) { return h ? h->sm_count : TKBO_ERR_INVALID; }
int64_t tkbo_launch_count(tkbo_handle_t h) { return h ? h->launches : TKBO_ERR_INVALID; }
int tkbo_debug_set_trace(tkbo_handle_t h, uint64_t* trace, int cap) {
  if (!h || cap < 0) return TKBO_ERR_INVALID;
  h->trace = reinterpret_cast<unsigned long long*>(trace);
  h->trace_cap = trace ? cap : 0;
  return TKBO_OK;
}

int tkbo_rff_basis_create(tkbo_handle_t h, tkbo_basis_t* out, int D, int d, int S) {
  TKBO_ENTER(h);
  return basis_create(h, out, D, d, S);
}
int tkbo_rff_basis_destroy(tkbo_handle_t h, tkbo_basis_t basis) {
  TKBO_ENTER(h);
  return basis_destroy(h, basis);
}
int tkbo_rff_basis_set(tkbo_handle_t h, tkbo_basis_t basis, const double* W, const double* b, const double* Theta,
                       int theta_is_transposed, void* stream) {
  TKBO_ENTER(h);
  return basis_set(h, basis, W, b, Theta, theta_is_transposed, as_stream(stream));
}
int tkbo_rff_basis_config(tkbo_handle_t h, tkbo_basis_t basis, int64_t N, int32_t* cfg) {
  TKBO_ENTER(h);
  return basis_config(h, basis, N, cfg);
}
int tkbo_rff_argmax(tkbo_handle_t h, tkbo_basis_t basis, const float* X, int64_t N, int64_t idx_offset, float* out_val,
                    int64_t* out_idx, void* stream) {
  TKBO_ENTER(h);
  return rff_argmax(h, basis, X, N, idx_offset, out_val, out_idx, as_stream(stream));
}
int tkbo_rff_argmax_key(tkbo_handle_t h, tkbo_basis_t basis, const float* X, int64_t N, int64_t idx_offset,
                        int64_t* out_key, void* stream) {
  TKBO_ENTER(h);
  return rff_argmax_key(h, basis, X, N, idx_offset, out_key, as_stream(stream));
}
int tkbo_unpack_argmax_key(tkbo_handle_t h, const int64_t* key, int S, float* out_val, int64_t* out_idx, void* stream) {
  TKBO_ENTER(h);
  return unpack_argmax_key(h, key, S, out_val, out_idx, as_stream(stream));
}
int tkbo_rff_eval(tkbo_handle_t h, tkbo_basis_t basis, const float* X, int64_t N, float* out_F, int64_t ldf, void* stream) {
  TKBO_ENTER(h);
  return rff_eval(h, basis, X, N, out_F, ldf, as_stream(stream));
}
int tkbo_merge_argmax(tkbo_handle_t h, const float* vals, const int64_t* idx, int G, int S, float* out_val,
                      int64_t* out_idx, void* stream) {
  TKBO_ENTER(h);
  return merge_argmax(h, vals, idx, G, S, out_val, out_idx, as_stream(stream));
}

int tkbo_rff_argmax_host(tkbo_handle_t h, const double* W, const double* b, const double* Theta, int D, int d, int S,
                         const float* X, int64_t N, float* out_val, int64_t* out_idx) {
  TKBO_ENTER(h);
  if (!W || !b || !Theta || !X || !out_val || !out_idx) { set_error("null pointer"); return TKBO_ERR_INVALID; }
  if (N < 1) { set_error("N must be >= 1"); return TKBO_ERR_INVALID; }
  tkbo_basis_t bs = nullptr;
  int rc = basis_create(h, &bs, D, d, S);
  if (rc != TKBO_OK) return rc;
  double *dW = nullptr, *db = nullptr, *dT = nullptr;
  float *dX = nullptr, *dval = nullptr;
  int64_t* didx = nullptr;
  cudaStream_t st = nullptr;
  cudaError_t e = cudaSuccess;
  e = e ? e : cudaMalloc(&dW, sizeof(double) * D * d);
  e = e ? e : cudaMalloc(&db, sizeof(double) * D);
  e = e ? e : cudaMalloc(&dT, sizeof(double) * D * S);
  e = e ? e : cudaMalloc(&dX, sizeof(float) * N * d);
  e = e ? e : cudaMalloc(&dval, sizeof(float) * S);
  e = e ? e : cudaMalloc(&didx, sizeof(int64_t) * S);
  e = e ? e : cudaMemcpyAsync(dW, W, sizeof(double) * D * d, cudaMemcpyHostToDevice, st);
  e = e ? e : cudaMemcpyAsync(db, b, sizeof(double) * D, cudaMemcpyHostToDevice, st);
  e = e ? e : cudaMemcpyAsync(dT, Theta, sizeof(double) * D * S, cudaMemcpyHostToDevice, st);
  e = e ? e : cudaMemcpyAsync(dX, X, sizeof(float) * N * d, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) {
    rc = basis_set(h, bs, dW, db, dT, 0, st);
    if (rc == TKBO_OK) rc = rff_argmax(h, bs, dX, N, 0, dval, didx, st);
    e = e ? e : cudaMemcpyAsync(out_val, dval, sizeof(float) * S, cudaMemcpyDeviceToHost, st);
    e = e ? e : cudaMemcpyAsync(out_idx, didx, sizeof(int64_t) * S, cudaMemcpyDeviceToHost, st);
    e = e ? e : cudaStreamSynchronize(st);
  }
  cudaFree(dW); cudaFree(db); cudaFree(dT); cudaFree(dX); cudaFree(dval); cudaFree(didx);
  basis_destroy(h, bs);
  if (e != cudaSuccess) { set_error(std::string("tkbo_rff_argmax_host: ") + cudaGetErrorString(e)); return TKBO_ERR_CUDA; }
  return rc;
}

int tkbo_rff_eval_batched(tkbo_handle_t h, const double* X, const double* W, const double* b, const double* theta,
                          int count, int n, int D, int d, double* out_f, int64_t* out_idx, void* stream) {
  TKBO_ENTER(h);
  return rff_eval_batched(h, X, W, b, theta, count, n, D, d, out_f, out_idx, as_stream(stream));
}
int tkbo_rff_features_batched(tkbo_handle_t h, const double* X, const double* W, const double* b, int count, int n,
                              int D, int d, double* out_phi, void* stream) {
  TKBO_ENTER(h);
  return rff_features_batched(h, X, W, b, count, n, D, d, out_phi, as_stream(stream));
}
int tkbo_rff_ascent_batched(tkbo_handle_t h, double* X, const double* W, const double* b, const double* theta, int count,
                            int n, int D, int d, double min_val, double max_val, int num_steps, double lr, double eps,
                            double rel_tol, int* steps_run, void* stream) {
  TKBO_ENTER(h);
  return rff_ascent_batched(h, X, W, b, theta, count, n, D, d, min_val, max_val, num_steps, lr, eps, rel_tol, steps_run,
                            as_stream(stream));
}

int tkbo_mpes_topk(tkbo_handle_t h, const float* fchi, const int32_t* fstar_argmax, int S, int64_t Q, int C, int M,
                   const int32_t* perms, int P, int k, int mode, double* out_mi, void* stream) {
  TKBO_ENTER(h);
  return mpes_topk(h, fchi, fstar_argmax, S, Q, C, M, perms, P, k, mode, out_mi, as_stream(stream));
}
int tkbo_argmax_rows(tkbo_handle_t h, const float* fstar, int S, int M, int32_t* out, void* stream) {
  TKBO_ENTER(h);
  return argmax_rows(h, fstar, S, M, out, as_stream(stream));
}
int tkbo_soft_copeland(tkbo_handle_t h, const float* f, int n, float* out_score, int64_t* out_argmax, void* stream) {
  TKBO_ENTER(h);
  return soft_copeland(h, f, n, out_score, out_argmax, as_stream(stream));
}

}  // extern "C"
