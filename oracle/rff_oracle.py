"""float64 NumPy restatement of the reference's random-Fourier-feature sampler.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  "ff.py" below means
``/root/reference/fourier_features.py``.  Every function is deterministic given its
array inputs; random draws are injected (``rng`` / explicit ``z``) so parity is defined
on identical ``W, b, theta`` exactly as BASELINE.json's north star asks.
"""
from __future__ import annotations

import numpy as np

TWO_PI = 2.0 * np.pi


# --------------------------------------------------------------------------------------
# ff.py:7-21  fourier_features
# --------------------------------------------------------------------------------------
def fourier_features(X, W, b):
    """sqrt(2/D) * cos((W X^T + b)^T)           [ff.py:18-21]

    X (count, n, d), W (count, D, d), b (count, D, 1) -> (count, n, D), float64.
    No kernel-variance factor (ff.py:9-11).
    """
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    D = W.shape[1]
    WX_b = W @ np.swapaxes(X, -1, -2) + b                      # (count, D, n)   ff.py:20
    return np.sqrt(2.0 / D) * np.cos(np.swapaxes(WX_b, -1, -2))  # (count, n, D)  ff.py:21


# --------------------------------------------------------------------------------------
# ff.py:24-53  sample_fourier_features
# --------------------------------------------------------------------------------------
def product_kernel_lengthscales(per_dim_lengthscales, d):
    """ff.py:33-40: a Product kernel over d/2 objective dims duplicates each sub-kernel's
    lengthscale for the [x, x'] halves of a duel."""
    input_dims = d // 2
    ls = np.zeros(d)
    for i in range(input_dims):
        ls[i] = per_dim_lengthscales[i]
        ls[i + input_dims] = per_dim_lengthscales[i]
    return ls


def draw_basis(count, D, lengthscales, rng):
    """W ~ N(0, diag(l^-2)) (ff.py:44-47), b ~ U[0, 2pi) (ff.py:48-51).

    The NumPy generator stream is not TensorFlow's; parity is on identical (W, b)."""
    lengthscales = np.atleast_1d(np.asarray(lengthscales, dtype=np.float64))
    d = lengthscales.shape[0]
    W = rng.standard_normal((count, D, d)) / lengthscales
    b = rng.uniform(0.0, TWO_PI, size=(count, D, 1))
    return W, b


def sample_fourier_features(X, lengthscales, D=100, rng=None):
    """ff.py:24-53 with the kernel reduced to its lengthscale vector. -> (phi, W, b)."""
    rng = np.random.default_rng() if rng is None else rng
    X = np.asarray(X, dtype=np.float64)
    count, _, d = X.shape
    ls = np.broadcast_to(np.asarray(lengthscales, dtype=np.float64), (d,))
    W, b = draw_basis(count, D, ls, rng)
    return fourier_features(X, W, b), W, b


# --------------------------------------------------------------------------------------
# ff.py:56-87  sample_theta_variational
# --------------------------------------------------------------------------------------
def sample_theta_variational(phi, q_mu, q_sqrt, z=None, rng=None):
    """theta = pinv-regression of q(u) samples onto the train features.

    phi (count, n, D); q_mu (n, 1); q_sqrt (1, n, n) lower triangular -> (count, D, 1).
    q = q_sqrt z + q_mu, z ~ N(0, I)                                  [ff.py:68-74]
    n >= D: theta = (phi^T phi)^-1 phi^T q ; else phi^T (phi phi^T)^-1 q   [ff.py:80-84]
    Explicit inverse, no jitter, exactly like tf.linalg.inv; a singular matrix raises
    numpy.linalg.LinAlgError (the analogue of tf.errors.InvalidArgumentError).
    """
    phi = np.asarray(phi, dtype=np.float64)
    q_mu = np.asarray(q_mu, dtype=np.float64)
    q_sqrt = np.asarray(q_sqrt, dtype=np.float64)
    count, n, D = phi.shape
    if z is None:
        rng = np.random.default_rng() if rng is None else rng
        z = rng.standard_normal((count, n))
    z = np.asarray(z, dtype=np.float64).reshape(count, n)
    q_samples = q_sqrt @ z[:, :, None] + q_mu                        # (count, n, 1)  ff.py:74
    phiT = np.swapaxes(phi, -1, -2)                                  # (count, D, n)  ff.py:77
    if n >= D:
        transform = np.linalg.inv(phiT @ phi) @ phiT                 # ff.py:81
    else:
        transform = phiT @ np.linalg.inv(phi @ phiT)                 # ff.py:82
    return transform @ q_samples                                     # (count, D, 1)  ff.py:84


# --------------------------------------------------------------------------------------
# ff.py:90-107  sample_features_weights (retry-on-singular)
# --------------------------------------------------------------------------------------
def sample_features_weights(X, lengthscales, q_mu, q_sqrt, D, rng):
    """Resample (phi, W, b, theta) while the inversion fails; 100 failures -> ValueError("Failed")."""
    fail_count = 0
    while True:
        try:
            phi, W, b = sample_fourier_features(X, lengthscales, D, rng)
            theta = sample_theta_variational(phi, q_mu, q_sqrt, rng=rng)
            return phi, W, b, theta
        except np.linalg.LinAlgError:
            fail_count += 1
        if fail_count >= 100:
            raise ValueError("Failed")


# --------------------------------------------------------------------------------------
# ff.py:133-157  gradient ascent on the sampled functions (Keras RMSprop, rho = 0)
# --------------------------------------------------------------------------------------
def maximizer_objective(x_star, W, b, theta):
    """loss = -mean(Phi(x*) theta) over all (count, n_init) starts        [ff.py:133-135]"""
    return -np.mean(fourier_features(x_star, W, b) @ theta)


def maximizer_objective_grad(x_star, W, b, theta):
    """d loss / d x_star, shape (count, n_init, d).

    f = sqrt(2/D) sum_j theta_j cos(w_j.x + b_j)  =>  df/dx = -sqrt(2/D) sum_j theta_j sin(.) w_j,
    loss = -mean f  =>  dloss/dx = +sqrt(2/D)/(count*n_init) sum_j theta_j sin(.) w_j.
    """
    count, n_init, _ = x_star.shape
    D = W.shape[1]
    arg = x_star @ np.swapaxes(W, -1, -2) + np.swapaxes(b, -1, -2)   # (count, n_init, D)
    s = np.sin(arg) * np.swapaxes(theta, -1, -2)                     # theta_j sin(.)
    return np.sqrt(2.0 / D) / (count * n_init) * (s @ W)


def rmsprop_rho0_ascent(x0, W, b, theta, min_val, max_val, num_steps=3000,
                        lr=1e-3, eps=1e-7, rel_tol=1e-7):
    """ff.py:139-157.  tf.keras.optimizers.RMSprop(rho=0.0) in TF 2.1 (defaults lr=1e-3,
    momentum=0, epsilon=1e-7, centered=False):  rms <- g^2 ;  x <- x - lr * g / (sqrt(rms) + eps),
    then the Variable constraint clips to [min_val, max_val] (ff.py:145).  Stops when
    |(loss - prev) / prev| < 1e-7 on the global mean loss (ff.py:154).  Returns (x, steps_run).
    """
    x = np.array(x0, dtype=np.float64, copy=True)
    prev = maximizer_objective(x, W, b, theta)
    steps = 0
    for _ in range(num_steps):
        g = maximizer_objective_grad(x, W, b, theta)
        x = x - lr * g / (np.sqrt(g * g) + eps)
        x = np.clip(x, min_val, max_val)
        cur = maximizer_objective(x, W, b, theta)
        steps += 1
        if abs((cur - prev) / prev) < rel_tol:
            break
        prev = cur
    return x, steps


# --------------------------------------------------------------------------------------
# ff.py:162-167  final evaluation, per-sample argmax over n_init, gather
# --------------------------------------------------------------------------------------
def eval_argmax_gather(x_star, W, b, theta):
    """fvals (count, n_init) = squeeze(Phi(x*) theta); argmax axis 1 (lowest index on ties,
    tf.math.argmax / np.argmax semantics); maximizers = x_star[s, argmax_s].   [ff.py:162-167]"""
    fvals = np.squeeze(fourier_features(x_star, W, b) @ theta, axis=2)
    idx = np.argmax(fvals, axis=1)
    maximizers = np.asarray(x_star)[np.arange(fvals.shape[0]), idx]
    return fvals, idx, maximizers


# --------------------------------------------------------------------------------------
# Shared-basis form of ff.py:162-167 with n_init -> N (the fused-kernel semantics).
# One (W, b) for all S posterior weight samples; Theta (D, S).  Chunked over candidates so
# that neither Phi (N x D) nor F (N x S) is ever materialised in full.
# --------------------------------------------------------------------------------------
def rff_values_shared(X, W, b, Theta, chunk=8192):
    """F[n, s] = sqrt(2/D) sum_j cos(X[n].W[j] + b[j]) Theta[j, s]   (ff.py:20-21 then '@ theta')."""
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    Theta = np.asarray(Theta, dtype=np.float64)
    N, D = X.shape[0], W.shape[0]
    out = np.empty((N, Theta.shape[1]))
    scale = np.sqrt(2.0 / D)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        out[lo:hi] = (scale * np.cos(X[lo:hi] @ W.T + b)) @ Theta
    return out


def rff_argmax_shared(X, W, b, Theta, chunk=8192, idx_offset=0, return_absmax=False):
    """Per-sample (max value, lowest arg max) over all N candidates, plus the runner-up value
    (for tie accounting: an index mismatch only counts when the top-2 gap exceeds tolerance).
    Returns (val[S], idx[S] int64, second[S]) and, with return_absmax, max_n |F[n, s]| (the scale
    relative tolerances are measured against)."""
    X = np.asarray(X, dtype=np.float64)
    W = np.asarray(W, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    Theta = np.asarray(Theta, dtype=np.float64)
    N, D = X.shape[0], W.shape[0]
    S = Theta.shape[1]
    scale = np.sqrt(2.0 / D)
    best = np.full(S, -np.inf)
    second = np.full(S, -np.inf)
    best_idx = np.zeros(S, dtype=np.int64)
    absmax = np.zeros(S)
    cols = np.arange(S)
    for lo in range(0, N, chunk):
        hi = min(N, lo + chunk)
        F = (scale * np.cos(X[lo:hi] @ W.T + b)) @ Theta            # (chunk, S)
        absmax = np.maximum(absmax, np.abs(F).max(axis=0))
        i1 = np.argmax(F, axis=0)
        v1 = F[i1, cols]
        if hi - lo > 1:
            F[i1, cols] = -np.inf
            v2 = F.max(axis=0)
        else:
            v2 = np.full(S, -np.inf)
        better = v1 > best                                            # strict: keep lowest index
        second = np.where(better, np.maximum(best, v2), np.maximum(second, v1))
        best_idx = np.where(better, i1 + lo, best_idx)
        best = np.where(better, v1, best)
    if return_absmax:
        return best, best_idx + idx_offset, second, absmax
    return best, best_idx + idx_offset, second


def theta_shared_basis(phi_train, q_mu, q_sqrt, Z):
    """Shared-basis weight posterior: the ff.py:80-84 transform is computed once for the one
    basis and applied to S posterior draws  q_s = q_sqrt z_s + q_mu,  Z (S, n).  -> Theta (D, S).
    Column s equals sample_theta_variational(phi_train[None], ..., z=Z[s:s+1])[0, :, 0]."""
    phi = np.asarray(phi_train, dtype=np.float64)                     # (n, D)
    n, D = phi.shape
    Q = np.asarray(q_sqrt, dtype=np.float64).reshape(n, n) @ np.asarray(Z, dtype=np.float64).T \
        + np.asarray(q_mu, dtype=np.float64).reshape(n, 1)            # (n, S)
    if n >= D:
        T = np.linalg.inv(phi.T @ phi) @ phi.T
    else:
        T = phi.T @ np.linalg.inv(phi @ phi.T)
    return T @ Q
