"""Random-Fourier-feature posterior sampling — same call signatures as the reference's
``fourier_features.py`` (cited below as ff.py), arithmetic in libtkbo.so (sm_100a).

Reference-shaped functions (per-sample basis, float64 end to end):
    fourier_features(X, W, b)                     ff.py:7-21     -> tkbo_rff_features_batched
    sample_fourier_features(X, kernel, D=100)     ff.py:24-53    (host RNG for W, b)
    sample_theta_variational(phi, q_mu, q_sqrt)   ff.py:56-87    (n x n / D x D inverse: torch.linalg = cuSOLVER;
                                                                  once per BO iteration, not the hot path)
    sample_features_weights(X, model, D)          ff.py:90-107   (retry on singular, <= 100, ValueError("Failed"))
    sample_maximizers(X, count, n_init, D, model, min_val, max_val, num_steps=3000)   ff.py:110-169
        -> tkbo_rff_ascent_batched (on-device sign-gradient ascent, ff.py:139-157)
           + tkbo_rff_eval_batched (final evaluation / argmax / gather, ff.py:162-167)

B200 fast path (one basis shared by all posterior samples, so Phi.Theta is one dense GEMM):
    SharedBasis.sample(...) / .argmax(candidates) / .eval(candidates) / .maximizers(candidates)
        -> tkbo_rff_basis_set + tkbo_rff_argmax / tkbo_rff_eval  (fused feature map + tcgen05 + argmax)
    sample_maximizers(..., candidates=grid) uses it: the per-sample argmax over an N-point candidate set is
    exactly ff.py:162-167 with n_init -> N (grid precedent: acquisitions/pes.py:24-41, dts.py:71-85).

Inputs may be NumPy arrays, torch tensors or anything with ``.numpy()``.  Results that experiment scripts
consume with NumPy (``sample_maximizers``) are CPU float64 torch tensors (``.numpy()`` / ``np.asarray`` work,
like the TF eager tensors the reference returns); intermediate arrays stay on the GPU.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native as nat
from .model import kernel_lengthscales, to_numpy

_rng = np.random.default_rng()


def seed(s):
    """Seed the host generator used for W, b, z and the ascent starting points."""
    global _rng
    _rng = np.random.default_rng(s)


class SingularMatrixError(ArithmeticError):
    """Raised where the reference gets tf.errors.InvalidArgumentError from tf.linalg.inv (ff.py:99)."""


def _device(device=None):
    torch = nat.require_cuda()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    if device.type != "cuda":
        raise nat.TkboError("the RFF path runs on CUDA devices only")
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def _f64(a, device):
    torch = nat.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(to_numpy(a))).to(device)


def _f32(a, device):
    torch = nat.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32).contiguous()
    arr = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(device)


# ---------------------------------------------------------------------------------------------------
# reference-shaped API (per-sample basis)
# ---------------------------------------------------------------------------------------------------
def fourier_features(X, W, b, device=None):
    """sqrt(2/D) cos((W X^T + b)^T).  X (count, n, d), W (count, D, d), b (count, D, 1) -> (count, n, D).  [ff.py:7-21]"""
    torch = nat.require_cuda()
    dev = _device(device)
    X, W, b = _f64(X, dev), _f64(W, dev), _f64(b, dev)
    count, n, d = X.shape
    D = W.shape[1]
    assert W.shape == (count, D, d) and b.numel() == count * D, "fourier_features: shape mismatch"
    phi = torch.empty((count, n, D), dtype=torch.float64, device=dev)
    nat.check(nat.lib().tkbo_rff_features_batched(nat.handle(dev.index), nat.ptr(X), nat.ptr(W), nat.ptr(b.reshape(count, D)),
                                                  count, n, D, d, nat.ptr(phi), nat.current_stream_ptr(dev)),
              "tkbo_rff_features_batched")
    return phi


def _draw_basis(count, D, lengthscales, rng):
    """W ~ N(0, diag(l^-2)), b ~ U[0, 2pi)                                                    [ff.py:44-51]"""
    d = lengthscales.shape[0]
    W = rng.standard_normal((count, D, d)) / lengthscales
    b = rng.uniform(0.0, 2.0 * np.pi, size=(count, D, 1))
    return W, b


def sample_fourier_features(X, kernel, D=100, device=None):
    """Draw a basis per leading index of X and return (phi, W, b).                            [ff.py:24-53]"""
    dev = _device(device)
    Xh = X if hasattr(X, "shape") else np.asarray(X)
    count, d = int(Xh.shape[0]), int(Xh.shape[2])
    W, b = _draw_basis(count, D, kernel_lengthscales(kernel, d), _rng)
    Wd, bd = _f64(W, dev), _f64(b, dev)
    return fourier_features(X, Wd, bd, device=dev), Wd, bd


def _regression_transform(phi):
    """(phi^T phi)^-1 phi^T if n >= D else phi^T (phi phi^T)^-1 — explicit inverse, no jitter.  [ff.py:80-82]"""
    torch = nat.require_cuda()
    n, D = phi.shape[-2], phi.shape[-1]
    phiT = phi.transpose(-1, -2)
    gram = phiT @ phi if n >= D else phi @ phiT
    inv, info = torch.linalg.inv_ex(gram, check_errors=False)        # LU via cuSOLVER; info > 0 <=> zero pivot
    if bool((info != 0).any()) or not bool(torch.isfinite(inv).all()):
        raise SingularMatrixError("Input is not invertible.")        # the reference gets InvalidArgumentError here
    return inv @ phiT if n >= D else phiT @ inv


def sample_theta_variational(phi, q_mu, q_sqrt, z=None):
    """theta (count, D, 1) from q = q_sqrt z + q_mu, z ~ N(0, I).                              [ff.py:56-87]
    ``z`` (count, n) may be supplied for reproducibility (the reference draws it with TFP)."""
    dev = _device(phi.device if hasattr(phi, "device") and getattr(phi.device, "type", "") == "cuda" else None)
    phi = _f64(phi, dev)
    count, n, _ = phi.shape
    if z is None:
        z = _rng.standard_normal((count, n))
    z = _f64(z, dev).reshape(count, n)
    q_samples = _f64(q_sqrt, dev).reshape(1, n, n) @ z[:, :, None] + _f64(q_mu, dev).reshape(n, 1)   # ff.py:74
    return _regression_transform(phi) @ q_samples                                                       # ff.py:84


def sample_features_weights(X, model, D, device=None):
    """(phi, W, b, theta) with resampling while the inverse fails; 100 failures -> ValueError("Failed").  [ff.py:90-107]"""
    fail_count = 0
    while True:
        try:
            phi, W, b = sample_fourier_features(X, model.kernel, D, device=device)
            theta = sample_theta_variational(phi, model.q_mu, model.q_sqrt)
            return phi, W, b, theta
        except SingularMatrixError as err:
            print(err)
            print("Resampling phi, W, b, theta")
            fail_count += 1
        if fail_count >= 100:
            print("Retry limit exceeded")
            raise ValueError("Failed")


def eval_argmax_gather(x_star, W, b, theta, device=None):
    """fvals (count, n_init), argmax over n_init (lowest index), gathered maximizers (count, d).  [ff.py:162-167]"""
    torch = nat.require_cuda()
    dev = _device(device)
    x_star, W, b, theta = _f64(x_star, dev), _f64(W, dev), _f64(b, dev), _f64(theta, dev)
    count, n_init, d = x_star.shape
    D = W.shape[1]
    fvals = torch.empty((count, n_init), dtype=torch.float64, device=dev)
    idx = torch.empty(count, dtype=torch.int64, device=dev)
    nat.check(nat.lib().tkbo_rff_eval_batched(nat.handle(dev.index), nat.ptr(x_star), nat.ptr(W), nat.ptr(b.reshape(count, D)),
                                              nat.ptr(theta.reshape(count, D)), count, n_init, D, d, nat.ptr(fvals),
                                              nat.ptr(idx), nat.current_stream_ptr(dev)), "tkbo_rff_eval_batched")
    maximizers = x_star[torch.arange(count, device=dev), idx]
    return fvals, idx, maximizers


def ascend(x_star, W, b, theta, min_val, max_val, num_steps=3000, lr=1e-3, eps=1e-7, rel_tol=1e-7, device=None):
    """On-device restatement of the optimisation loop ff.py:139-157: tf.keras RMSprop(rho=0) on
    loss = -mean(Phi(x) theta), clip constraint, early stop on |dloss/loss| < 1e-7.  Returns (x, steps_run)."""
    dev = _device(device)
    x = _f64(x_star, dev).clone()
    W, b, theta = _f64(W, dev), _f64(b, dev), _f64(theta, dev)
    count, n_init, d = x.shape
    D = W.shape[1]
    steps = ctypes.c_int(0)
    nat.check(nat.lib().tkbo_rff_ascent_batched(nat.handle(dev.index), nat.ptr(x), nat.ptr(W), nat.ptr(b.reshape(count, D)),
                                                nat.ptr(theta.reshape(count, D)), count, n_init, D, d, float(min_val),
                                                float(max_val), int(num_steps), float(lr), float(eps), float(rel_tol),
                                                ctypes.byref(steps), nat.current_stream_ptr(dev)), "tkbo_rff_ascent_batched")
    return x, steps.value


def sample_maximizers(X, count, n_init, D, model, min_val, max_val, num_steps=3000, *, candidates=None,
                      device=None, return_info=False):
    """Samples from the posterior over the global maximiser (Shah & Ghahramani 2015).        [ff.py:110-169]

    Default (``candidates=None``) follows the reference step for step: a basis and theta per sample, n_init uniform
    starts per sample, sign-gradient ascent for <= num_steps, then argmax over the starts -> (count, d).

    With ``candidates`` (N, d): grid strategy on the fused kernel — one shared basis, Theta (D, count), and the
    per-sample argmax over all N candidates (ff.py:162-167 with n_init -> N); n_init / num_steps are ignored.
    """
    torch = nat.require_cuda()
    dev = _device(device)
    Xh = to_numpy(X)
    d = Xh.shape[1]
    if candidates is not None:
        basis = SharedBasis.sample(Xh, model, D=D, count=count, device=dev)
        val, idx = basis.argmax(candidates)
        cand = candidates if isinstance(candidates, torch.Tensor) else torch.from_numpy(np.asarray(candidates))
        maxi = cand.to(dev)[idx].to(torch.float64).cpu()
        return (maxi, {"values": val.cpu(), "indices": idx.cpu(), "basis": basis}) if return_info else maxi

    Xt = np.tile(Xh[None], (count, 1, 1))                                   # ff.py:129
    phi, W, b, theta = sample_features_weights(Xt, model, D, device=dev)    # ff.py:131
    x0 = _rng.uniform(min_val, max_val, size=(count, n_init, d))            # ff.py:141-144
    x_star, steps = ascend(x0, W, b, theta, min_val, max_val, num_steps=num_steps, device=dev)
    fvals, idx, maximizers = eval_argmax_gather(x_star, W, b, theta, device=dev)
    out = maximizers.cpu()
    if return_info:
        return out, {"steps": steps, "fvals": fvals.cpu(), "indices": idx.cpu(), "x_star": x_star.cpu(),
                     "W": W, "b": b, "theta": theta}
    return out


# ---------------------------------------------------------------------------------------------------
# shared-basis fast path
# ---------------------------------------------------------------------------------------------------
def argmax_host(W, b, Theta, X, device=0, out_val=None, out_idx=None, peer=None, idx_offset=0, precision="auto"):
    """End-to-end host-buffer call (C ABI ``tkbo_rff_argmax_host``): W (D, d), b (D,), Theta (D, S) float64 and the
    candidates X (N, d) float32 are HOST arrays (NumPy, or CPU torch tensors - with pinned memory the single kernel
    launch consumes the candidate chunks while their H2D copies are still landing); returns (val[S] float32,
    idx[S] int64) NumPy arrays.
    Device staging and the basis layout are cached on the handle between calls of the same shape.
    ``peer`` (a distributed.PeerMerge) + ``idx_offset``: X is this rank's shard of a sharded candidate set and the results
    are the GLOBAL ones (``tkbo_rff_argmax_host_peer``: the same single launch also delivers the keys to the peers and
    merges; every rank of the group must make the call).  ``precision``: "auto" / "fp32" = the fp32-emulating form,
    "fast" = the e4m3 correction form (see SharedBasis; the caller vouches for ||theta'|| / max|f|)."""
    def host(a, dtype):
        if hasattr(a, "data_ptr"):
            assert a.device.type == "cpu" and a.is_contiguous()
            return a, a.data_ptr(), tuple(a.shape)
        a = np.ascontiguousarray(a, dtype=dtype)
        return a, a.ctypes.data, a.shape
    Wk, Wp, Ws = host(W, np.float64)
    bk, bp, bs_ = host(b, np.float64)
    Tk, Tp, Ts = host(Theta, np.float64)
    Xk, Xp, Xs = host(X, np.float32)
    D, d = Ws
    S = Ts[1]
    assert Ts[0] == D and int(np.prod(bs_)) == D and Xs[1] == d
    val = np.empty(S, dtype=np.float32) if out_val is None else out_val
    idx = np.empty(S, dtype=np.int64) if out_idx is None else out_idx
    vp = val.data_ptr() if hasattr(val, "data_ptr") else val.ctypes.data
    ip = idx.data_ptr() if hasattr(idx, "data_ptr") else idx.ctypes.data
    dev = device if isinstance(device, int) else _device(device).index
    if peer is None:
        nat.check(nat.lib().tkbo_rff_argmax_host(nat.handle(dev), ctypes.c_void_p(Wp), ctypes.c_void_p(bp), ctypes.c_void_p(Tp),
                                                 int(D), int(d), int(S), ctypes.c_void_p(Xp), int(Xs[0]),
                                                 SharedBasis.PRECISIONS[precision], ctypes.c_void_p(vp),
                                                 ctypes.c_void_p(ip)), "tkbo_rff_argmax_host")
    else:
        peer.epoch += 1
        nat.check(nat.lib().tkbo_rff_argmax_host_peer(nat.handle(dev), ctypes.c_void_p(Wp), ctypes.c_void_p(bp),
                                                      ctypes.c_void_p(Tp), int(D), int(d), int(S), ctypes.c_void_p(Xp),
                                                      int(Xs[0]), SharedBasis.PRECISIONS[precision], int(idx_offset), peer._ptrs,
                                                      peer.world, peer.rank,
                                                      peer.epoch, ctypes.c_void_p(vp), ctypes.c_void_p(ip)),
                  "tkbo_rff_argmax_host_peer")
    return val, idx


class SharedBasis:
    """One RFF basis (W (D, d), b (D,)) shared by S posterior weight samples Theta (D, S), resident on one GPU in
    the fused kernel's operand layout.  ``argmax`` / ``eval`` stream any number of candidates through it."""

    PRECISIONS = {"auto": 0, "fp32": 1, "fast": 2}
    FAST_ERROR_PER_NORM = 1e-4      # |f_fast - f_fp64| <= this * ||sqrt(2/D) theta_s||_2 (measured <= ~1.5e-5; see error_bound)
    FAST_BUDGET = 0.3e-4            # share of the path's 1e-4 tolerance the fast form may use, relative to max|f_s|

    def __init__(self, W, b, Theta, device=None, precision="auto", f_scale=None):
        """``precision``: "fp32" = three fp16 tensor-core products per term (fp32 emulation, ~2e-6 of max|f|); "fast" =
        fp16 hi*hi + one e4m3 product for both correction terms (two thirds of the tensor work; error proportional to
        ||sqrt(2/D) theta_s||, ~3e-5 of max|f| on a well-posed regression); "auto" = fast only if it is provably inside
        the tolerance: the caller passes ``f_scale`` (S,) = a lower bound of max|f_s| (``sample`` uses the q-samples the
        weights interpolate, ff.py:74) and FAST_ERROR_PER_NORM * ||theta'_s|| <= FAST_BUDGET * f_scale_s holds for every
        sample, and the sample count is in the tensor-bound regime (> 96); fp32 otherwise."""
        torch = nat.require_cuda()
        self.device = _device(device)
        self.W, self.b, self.Theta = _f64(W, self.device), _f64(b, self.device).reshape(-1), _f64(Theta, self.device)
        self.D, self.d = self.W.shape
        self.S = self.Theta.shape[1]
        assert self.Theta.shape[0] == self.D and self.b.shape[0] == self.D, "SharedBasis: shape mismatch"
        self._h = nat.handle(self.device.index)
        self._torch = torch
        self._requested = precision
        self._basis = ctypes.c_void_p()
        self.precision = self._choose_precision(precision, f_scale)
        self._create()

    def _choose_precision(self, precision, f_scale):
        if precision != "auto":
            assert precision in self.PRECISIONS, "precision must be 'auto', 'fp32' or 'fast'"
            return precision
        if f_scale is None or self.S <= 96:
            return "fp32"
        torch = self._torch
        fs = torch.as_tensor(np.asarray(to_numpy(f_scale)).reshape(-1)).to(self.device)
        norm = (self.Theta * float(np.sqrt(2.0 / self.D))).norm(dim=0)
        ok = bool((self.FAST_ERROR_PER_NORM * norm <= self.FAST_BUDGET * fs).all())
        return "fast" if ok else "fp32"

    def _create(self):
        if self._basis:
            nat.lib().tkbo_rff_basis_destroy(self._h, self._basis)
            self._basis = ctypes.c_void_p()
        nat.check(nat.lib().tkbo_rff_basis_create_ex(self._h, ctypes.byref(self._basis), self.D, self.d, self.S,
                                                     self.PRECISIONS[self.precision]), "tkbo_rff_basis_create_ex")
        nat.check(nat.lib().tkbo_rff_basis_set(self._h, self._basis, nat.ptr(self.W), nat.ptr(self.b), nat.ptr(self.Theta), 0,
                                               nat.current_stream_ptr(self.device)), "tkbo_rff_basis_set")

    def set(self, W, b, Theta, f_scale=None):
        """Load a new draw of the same shape (next BO iteration) without reallocating the device layout (with
        precision="auto" the form is re-decided for the new weights; a change of form rebuilds the layout)."""
        W, b, Theta = _f64(W, self.device), _f64(b, self.device).reshape(-1), _f64(Theta, self.device)
        assert W.shape == (self.D, self.d) and b.shape == (self.D,) and Theta.shape == (self.D, self.S)
        self.W, self.b, self.Theta = W, b, Theta
        want = self._choose_precision(self._requested, f_scale)
        if want != self.precision:
            self.precision = want
            self._create()
            return self
        nat.check(nat.lib().tkbo_rff_basis_set(self._h, self._basis, nat.ptr(W), nat.ptr(b), nat.ptr(Theta), 0,
                                               nat.current_stream_ptr(self.device)), "tkbo_rff_basis_set")
        return self

    def __del__(self):
        try:
            if getattr(self, "_basis", None):
                nat.lib().tkbo_rff_basis_destroy(self._h, self._basis)
                self._basis = None
        except Exception:
            pass

    @classmethod
    def sample(cls, X_train, model, D, count, device=None, rng=None, precision="auto"):
        """Shared-basis analogue of sample_features_weights: draw (W, b) once, fit theta_s = T q_s for
        s < count with T the ff.py:80-82 transform on the training features; retry on singular (ff.py:92-105)."""
        dev = _device(device)
        rng = _rng if rng is None else rng
        Xh = to_numpy(X_train)
        n, d = Xh.shape
        ls = kernel_lengthscales(model.kernel, d)
        fail_count = 0
        while True:
            try:
                W, b = _draw_basis(1, D, ls, rng)
                phi = fourier_features(Xh[None], W, b, device=dev)[0]                     # (n, D)
                Z = _f64(rng.standard_normal((count, n)), dev)
                Q = _f64(model.q_sqrt, dev).reshape(n, n) @ Z.T + _f64(model.q_mu, dev).reshape(n, 1)   # (n, count)
                Theta = _regression_transform(phi) @ Q                                    # (D, count)
                # the weights interpolate the q-samples on the training inputs (ff.py:80-84), so max_i |q_is| is a lower
                # bound of max|f_s|: the scale the 1e-4 tolerance is measured against
                return cls(W[0], b[0, :, 0], Theta, device=dev, precision=precision, f_scale=Q.abs().amax(dim=0))
            except SingularMatrixError as err:
                print(err)
                print("Resampling phi, W, b, theta")
                fail_count += 1
            if fail_count >= 100:
                print("Retry limit exceeded")
                raise ValueError("Failed")

    def error_bound(self, rel_feature_error=None):
        """Per-sample absolute error bound (S,) of the fused kernel's values: the feature map is evaluated in
        fp32 (argument rounding, cos.approx, fp16 hi/lo split: each cos value is good to ~1e-6, measured
        <= 4e-6 for |x.w + b| up to ~50), so |f_gpu - f_fp64| <= rel_feature_error * ||sqrt(2/D) theta_s||_2 * sqrt(D)
        in the worst case and ~rel_feature_error * ||sqrt(2/D) theta_s||_2 in practice (independent errors).
        A well-posed posterior has ||sqrt(2/D) theta_s|| ~ max|f_s|, i.e. ~1e-6 relative; an ill-posed
        regression (ff.py:80-82 inverts phi phi^T without jitter) can inflate it by orders of magnitude.
        With the e4m3 correction form (precision "fast") the per-product error is ~2^-12 * 2^-4 instead: FAST_ERROR_PER_NORM
        (1e-4; measured <= 1.5e-5 per unit norm, the rest is margin for the tail over 2^24 x S values)."""
        if rel_feature_error is None:
            rel_feature_error = self.FAST_ERROR_PER_NORM if self.config(1)["e4m3_corrections"] else 4e-6
        return rel_feature_error * (self.Theta * float(np.sqrt(2.0 / self.D))).norm(dim=0)

    def config(self, N):
        cfg = (ctypes.c_int32 * 16)()
        nat.check(nat.lib().tkbo_rff_basis_config(self._h, self._basis, int(N), cfg), "tkbo_rff_basis_config")
        keys = ["S_chunk", "n_chunks", "n_slabs", "stages", "acc_bufs", "d_template", "smem_bytes", "grid",
                "tmem_ring", "threads", "row_bufs", "producer_groups", "proj_warps", "e4m3_corrections", "w_stages", "producers_share_slab"]
        out = dict(zip(keys, list(cfg)))
        out["cta_pairs"] = out["producers_share_slab"] >> 1           # cfg[15] = coop + 2 * pair
        out["producers_share_slab"] &= 1
        return out

    def two_pass_items(self):
        """(items the full-precision pass of the last argmax covered, items in all), or None if it ran as a single pass.
        Large argmax calls run a hi*hi-only pass over everything and the full contraction only where the maximum can be
        (csrc/rff.cu: launch_argmax); the results are the same bits either way.  Synchronises the device."""
        self._torch.cuda.synchronize(self.device)
        a, b = ctypes.c_int(-1), ctypes.c_int(-1)
        nat.check(nat.lib().tkbo_two_pass_items(self._h, ctypes.byref(a), ctypes.byref(b)), "tkbo_two_pass_items")
        return None if a.value < 0 else (a.value, b.value)

    def _candidates(self, X):
        X = _f32(X, self.device)
        assert X.dim() == 2 and X.shape[1] == self.d, "candidates must be (N, d)"
        return X

    def argmax(self, X, idx_offset=0):
        """(val[S] float32, idx[S] int64), CUDA tensors: max and lowest arg max over the N candidates."""
        torch = self._torch
        X = self._candidates(X)
        val = torch.empty(self.S, dtype=torch.float32, device=self.device)
        idx = torch.empty(self.S, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().tkbo_rff_argmax(self._h, self._basis, nat.ptr(X), X.shape[0], int(idx_offset), nat.ptr(val),
                                            nat.ptr(idx), nat.current_stream_ptr(self.device)), "tkbo_rff_argmax")
        return val, idx

    def argmax_key(self, X, idx_offset=0):
        """(S,) int64 CUDA tensor of merge keys: integer max == largest value then lowest global index, so shards
        held by different ranks merge with a single all_reduce(MAX) (see distributed.allreduce_argmax)."""
        torch = self._torch
        X = self._candidates(X)
        key = torch.empty(self.S, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().tkbo_rff_argmax_key(self._h, self._basis, nat.ptr(X), X.shape[0], int(idx_offset), nat.ptr(key),
                                                nat.current_stream_ptr(self.device)), "tkbo_rff_argmax_key")
        return key

    def argmax_peer(self, X, idx_offset, peer_ptrs, rank, epoch, out_val=None, out_idx=None):
        """Fused kernel + merge push: this rank's S keys go straight into slot (epoch & 1, rank) of every rank's peer
        buffer (``peer_ptrs`` = ctypes array of the buffers as mapped here), followed by a release of flag[rank] = epoch;
        with ``out_val`` / ``out_idx`` the same launch's last CTA also merges all ranks' keys into them;
        see distributed.PeerMerge, include/tkbo.h tkbo_rff_argmax_peer."""
        X = self._candidates(X)
        nat.check(nat.lib().tkbo_rff_argmax_peer(self._h, self._basis, nat.ptr(X), X.shape[0], int(idx_offset), peer_ptrs,
                                                 len(peer_ptrs), int(rank), int(epoch), nat.ptr(out_val), nat.ptr(out_idx),
                                                 nat.current_stream_ptr(self.device)),
                  "tkbo_rff_argmax_peer")

    def unpack_key(self, key):
        """(val[S] float32, idx[S] int64) from merge keys (idx = -1 where no candidate was seen)."""
        torch = self._torch
        key = key.contiguous()
        val = torch.empty(key.shape[0], dtype=torch.float32, device=self.device)
        idx = torch.empty(key.shape[0], dtype=torch.int64, device=self.device)
        nat.check(nat.lib().tkbo_unpack_argmax_key(self._h, nat.ptr(key), key.shape[0], nat.ptr(val), nat.ptr(idx),
                                                   nat.current_stream_ptr(self.device)), "tkbo_unpack_argmax_key")
        return val, idx

    def duel_copeland(self, X_base, scores=True):
        """Dueling-Thompson soft-Copeland for all S samples at once: the basis lives on duels [x, x'] (d = 2 d_obj),
        X_base (n, d_obj) are the base points; returns (score[S, n] float32 or None, winner[S] int64) CUDA tensors with
        score[s, i] = mean_j sigmoid(f_s([x_i, x_j])) (dts.py:94-96).  The n^2 duels are generated inside the kernel."""
        torch = self._torch
        Xb = _f32(X_base, self.device)
        assert Xb.dim() == 2 and 2 * Xb.shape[1] == self.d, "base points must be (n, d / 2)"
        n = Xb.shape[0]
        score = torch.empty((self.S, n), dtype=torch.float32, device=self.device) if scores else None
        win = torch.empty(self.S, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().tkbo_rff_duel_copeland(self._h, self._basis, nat.ptr(Xb), n,
                                                   nat.ptr(score) if scores else None, nat.ptr(win),
                                                   nat.current_stream_ptr(self.device)), "tkbo_rff_duel_copeland")
        return score, win

    def duel_copeland_partial(self, X_base, j_begin, j_end, fix=None):
        """Shard of duel_copeland: adds the j in [j_begin, j_end) part of the fixed-point sigmoid sums into ``fix``
        ((S, n) int64 CUDA tensor, created zeroed when None) and returns it; see distributed.sharded_duel_copeland."""
        torch = self._torch
        Xb = _f32(X_base, self.device)
        assert Xb.dim() == 2 and 2 * Xb.shape[1] == self.d, "base points must be (n, d / 2)"
        n = Xb.shape[0]
        if fix is None:
            fix = torch.zeros((self.S, n), dtype=torch.int64, device=self.device)
        assert fix.shape == (self.S, n) and fix.dtype == torch.int64 and fix.is_contiguous()
        nat.check(nat.lib().tkbo_rff_duel_copeland_partial(self._h, self._basis, nat.ptr(Xb), n, int(j_begin), int(j_end),
                                                           nat.ptr(fix), nat.current_stream_ptr(self.device)),
                  "tkbo_rff_duel_copeland_partial")
        return fix

    def copeland_finish(self, fix, scores=True):
        """(score[S, n] float32 or None, winner[S] int64) from complete fixed-point sums."""
        torch = self._torch
        S, n = fix.shape
        score = torch.empty((S, n), dtype=torch.float32, device=self.device) if scores else None
        win = torch.empty(S, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().tkbo_copeland_finish(self._h, nat.ptr(fix), int(S), int(n), nat.ptr(score) if scores else None,
                                                 nat.ptr(win), nat.current_stream_ptr(self.device)), "tkbo_copeland_finish")
        return score, win

    def eval(self, X, out=None):
        """(S, N) float32 CUDA tensor, F[s, n] = f_s(X[n])."""
        torch = self._torch
        X = self._candidates(X)
        N = X.shape[0]
        if out is None:
            out = torch.empty((self.S, N), dtype=torch.float32, device=self.device)
        assert out.shape == (self.S, N) and out.is_contiguous() and out.dtype == torch.float32
        nat.check(nat.lib().tkbo_rff_eval(self._h, self._basis, nat.ptr(X), N, nat.ptr(out), N,
                                          nat.current_stream_ptr(self.device)), "tkbo_rff_eval")
        return out

    def maximizers(self, X):
        """Candidates attaining each sample's maximum, (S, d) CPU float64 (the return type of sample_maximizers)."""
        Xc = self._candidates(X)
        _, idx = self.argmax(Xc)
        return Xc[idx].to(self._torch.float64).cpu()
