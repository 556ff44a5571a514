"""Duck-typed model objects for the hot path.

The reference reads four things from a GPflow 2.0.0-rc1 ``SVGP`` (models/learning.py:327-352, whiten=False):
``model.kernel.lengthscale`` (or ``kernel.kernels[i].lengthscale`` for a Product kernel, ff.py:33-42),
``model.q_mu`` (n, 1), ``model.q_sqrt`` (1, n, n) and ``model.predict_f_samples(X, S)``.  Training that
model is outside this package's scope; these classes carry an already-fitted posterior so the sampler
and acquisitions can be driven without GPflow.  A real GPflow model works too: everything here is
accessed by attribute name only.
"""
from __future__ import annotations

import numpy as np

from . import _native


def to_numpy(a):
    """np.float64 array from numpy / torch / tf-like (``.numpy()``) inputs."""
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    elif hasattr(a, "numpy"):
        a = a.numpy()
    return np.asarray(a, dtype=np.float64)


class Parameter:
    """Minimal stand-in for a gpflow Parameter: ``.numpy()`` gives the constrained value."""

    def __init__(self, value):
        self.value = np.asarray(value, dtype=np.float64)

    def numpy(self):
        return self.value

    def __repr__(self):
        return f"Parameter({self.value!r})"


class RBF:
    """Squared-exponential kernel, ARD lengthscale(s); GPflow 2.0.0-rc1 spells the attribute ``lengthscale``."""

    def __init__(self, lengthscale=1.0, variance=1.0, active_dims=None):
        self.lengthscale = Parameter(lengthscale)
        self.variance = Parameter(variance)
        self.active_dims = active_dims

    @property
    def lengthscales(self):  # later GPflow spelling
        return self.lengthscale

    def K(self, X, X2=None):
        X = to_numpy(X)
        X2 = X if X2 is None else to_numpy(X2)
        if self.active_dims is not None:
            X, X2 = X[:, self.active_dims], X2[:, self.active_dims]
        ls = np.broadcast_to(self.lengthscale.numpy(), (X.shape[1],))
        d2 = (((X[:, None, :] - X2[None, :, :]) / ls) ** 2).sum(-1)
        return float(self.variance.numpy()) * np.exp(-0.5 * d2)


class Product:
    """Product of per-dimension kernels, as the DTS scripts build it (ff.py:33-40 reads ``.kernels``)."""

    def __init__(self, kernels):
        self.kernels = list(kernels)

    def K(self, X, X2=None):
        out = None
        for k in self.kernels:
            kk = k.K(X, X2)
            out = kk if out is None else out * kk
        return out


def is_product_kernel(kernel):
    """ff.py:33 tests ``type(kernel) == gpflow.kernels.base.Product``; duck-typed here."""
    return hasattr(kernel, "kernels") and not hasattr(kernel, "lengthscale") and not hasattr(kernel, "lengthscales")


def kernel_lengthscales(kernel, d):
    """Per-input-dimension lengthscale vector of length d, following ff.py:33-42."""
    def val(k):
        p = getattr(k, "lengthscale", None)
        if p is None:
            p = getattr(k, "lengthscales")
        return to_numpy(p)

    if is_product_kernel(kernel):
        input_dims = d // 2
        ls = np.zeros(d)
        for i in range(input_dims):
            v = float(np.asarray(val(kernel.kernels[i])).reshape(-1)[0])
            ls[i] = v
            ls[i + input_dims] = v
        return ls
    return np.broadcast_to(val(kernel), (d,)).astype(np.float64)


class PreferentialGP:
    """A fitted variational preferential-GP posterior: inducing inputs Z (n, d) with q(u) = N(q_mu, q_sqrt q_sqrt^T).

    ``predict_f_samples`` follows GPflow's non-whitened SVGP conditional (the algebra the reference restates at
    models/learning.py:132-159): mu = Kxz Kzz^-1 q_mu, Sigma = Kxx - Kxz Kzz^-1 Kzx + A S A^T with A = Kxz Kzz^-1,
    samples = mu + chol(Sigma + jitter I) eps, shape (S, N, 1).  It is O(N^3) host linear algebra meant for the
    reference's small query sizes; ``predict_f_samples_device`` is the scalable RFF path (K1 'store-F' kernel).
    """

    def __init__(self, kernel, inducing_points, q_mu, q_sqrt, jitter=1e-6, seed=None):
        self.kernel = kernel
        self.Z = to_numpy(inducing_points)
        self.q_mu = to_numpy(q_mu).reshape(-1, 1)
        n = self.q_mu.shape[0]
        self.q_sqrt = to_numpy(q_sqrt).reshape(1, n, n)
        self.jitter = jitter
        self.rng = np.random.default_rng(seed)

    # -- exact joint sampling (host, float64) ------------------------------------------------------
    def predict_f(self, X, full_cov=False):
        X = to_numpy(X)
        Kzz = self.kernel.K(self.Z) + self.jitter * np.eye(self.Z.shape[0])
        Kxz = self.kernel.K(X, self.Z)
        A = np.linalg.solve(Kzz, Kxz.T).T                               # Kxz Kzz^-1
        mu = A @ self.q_mu
        Sq = self.q_sqrt[0] @ self.q_sqrt[0].T
        cov = self.kernel.K(X) - A @ Kxz.T + A @ Sq @ A.T
        return (mu, cov) if full_cov else (mu, np.diag(cov)[:, None])

    def predict_f_samples(self, X, num_samples):
        import torch
        mu, cov = self.predict_f(X, full_cov=True)
        N = mu.shape[0]
        L = np.linalg.cholesky(cov + self.jitter * np.eye(N))
        eps = self.rng.standard_normal((int(num_samples), N))
        return torch.from_numpy((mu[:, 0][None, :] + eps @ L.T)[:, :, None])   # (S, N, 1), has .numpy()

    # -- RFF joint sampling on the GPU (float32 result, sample-major) ------------------------------
    def predict_f_samples_device(self, X, num_samples, D=1024, device=None):
        """(S, N) float32 CUDA tensor of joint posterior samples at X drawn with one shared RFF basis:
        theta_s fitted on the inducing inputs (ff.py:56-87), f_s(X) = Phi(X) theta_s evaluated by the
        fused kernel's store-F epilogue.  Never builds an N x N covariance."""
        from . import fourier_features as ff
        basis = ff.SharedBasis.sample(self.Z, self, D=D, count=int(num_samples), device=device, rng=self.rng)
        return basis.eval(X)


def synthetic_model(n, d, low, high, lengthscale, seed=0, product=False):
    """The synthetic fitted posterior used by tests/bench: Z ~ U[low, high]^d,
    q_mu = L z, q_sqrt = 0.3 L, L = chol(K_SE(Z) + 1e-6 I)."""
    rng = np.random.default_rng(seed)
    Z = rng.uniform(low, high, size=(n, d))
    ls = np.broadcast_to(np.asarray(lengthscale, dtype=np.float64), (d,))
    if product:
        half = d // 2
        kernel = Product([RBF(ls[i], active_dims=[i, i + half]) for i in range(half)])
    else:
        kernel = RBF(ls)
    Kzz = np.exp(-0.5 * (((Z[:, None, :] - Z[None, :, :]) / ls) ** 2).sum(-1))
    L = np.linalg.cholesky(Kzz + 1e-6 * np.eye(n))
    q_mu = L @ rng.standard_normal((n, 1))
    return PreferentialGP(kernel, Z, q_mu, (0.3 * L)[None], seed=seed + 1)


__all__ = ["Parameter", "RBF", "Product", "PreferentialGP", "synthetic_model", "kernel_lengthscales",
           "is_product_kernel", "to_numpy", "_native"]
