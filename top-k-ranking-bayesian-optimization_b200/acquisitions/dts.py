"""Dueling-Thompson sampling (Gonzalez et al. 2017) — the hot-path half of the reference's acquisitions/dts.py:
``sample_f`` (dts.py:71-85) and ``soft_copeland_maximizer`` (dts.py:88-97), plus the two grid helpers the
scripts call (dts.py:12-45) and the Gauss-Hermite variance helpers (dts.py:48-68; small host-side torch code).
"""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _native as nat
from .. import fourier_features
from ..model import to_numpy
from . import _common


def uniform_grid(input_dims, num_discrete_per_dim, low, high):
    """All grid points, first dimension fastest.                                          [dts.py:12-28]"""
    pts = np.linspace(low, high, num_discrete_per_dim)
    idx = np.arange(num_discrete_per_dim ** input_dims)
    out = np.zeros([idx.shape[0], input_dims])
    for dim in range(input_dims):
        out[:, dim] = pts[(idx // (num_discrete_per_dim ** dim)) % num_discrete_per_dim]
    return out


def combinations(points):
    """All ordered pairs [x_i, x_j], i-major: (n^2, 2d).                                  [dts.py:31-45]"""
    points = to_numpy(points)
    n, d = points.shape
    out = np.empty((n, n, 2 * d))
    out[:, :, :d] = points[:, None, :]
    out[:, :, d:] = points[None, :, :]
    return out.reshape(n * n, 2 * d)


def logistic(x):
    """[dts.py:48-49]"""
    import torch
    return 1 / (1 + torch.exp(-torch.as_tensor(x)))


def logistic_square(x):
    """[dts.py:52-53]"""
    return logistic(x) ** 2


def _gauss_hermite_1d(func, means, variances, H=50):
    """E_{f ~ N(mean, var)}[func(f)] with H-point Gauss-Hermite quadrature: what gpflow.quadrature.mvnquad(func,
    means, variances, H, Din=1) computes.  means (N, 1), variances (N, 1, 1) or (N, 1) -> (N, 1) float64."""
    import torch
    xn, wn = np.polynomial.hermite.hermgauss(H)
    mu = torch.as_tensor(to_numpy(means), dtype=torch.float64).reshape(-1, 1)
    var = torch.as_tensor(to_numpy(variances), dtype=torch.float64).reshape(-1, 1)
    X = mu + torch.sqrt(2.0 * var) * torch.from_numpy(xn)[None, :]              # (N, H)
    return (func(X) * torch.from_numpy(wn / np.sqrt(np.pi))[None, :]).sum(dim=1, keepdim=True)


def variance_integral(means, variances):
    """E[sigmoid(f)^2]                                                                    [dts.py:56-57]"""
    return _gauss_hermite_1d(logistic_square, means, variances, H=50)


def expected_integral(means, variances):
    """E[sigmoid(f)]                                                                      [dts.py:60-61]"""
    return _gauss_hermite_1d(logistic, means, variances, H=50)


def variance_logistic_f(m, x):
    """Var[sigmoid(f(x))] under the posterior marginals of ``m.predict_f`` -> (N, 1).        [dts.py:64-68]"""
    means, variances = m.predict_f(to_numpy(x))
    return variance_integral(means, variances) - expected_integral(means, variances) ** 2


def sample_f(m, query_points, combs, D=1000, *, device=None, device_output=False):
    """One continuous-Thompson sample of f on every duel in ``combs``.                     [dts.py:71-85]
    W, b drawn for the (Product) kernel (ff.py:33-51), theta fitted on Phi(query_points) (dts.py:83-84),
    f = Phi(combs) theta through the fused kernel's store-F epilogue -> (n^2, 1).
    Returns a CPU float64 tensor (``.numpy()`` works as on the reference's TF tensor) or, with
    ``device_output=True``, the float32 CUDA tensor to hand to ``soft_copeland_maximizer`` without a copy."""
    basis = fourier_features.SharedBasis.sample(to_numpy(query_points), m, D=D, count=1, device=device)
    F = basis.eval(combs)                       # (1, n^2) float32 on the GPU
    f = F.reshape(-1, 1)
    return f if device_output else f.double().cpu()


def soft_copeland_maximizer(f_vals, discrete_space, *, device=None, return_scores=False):
    """argmax_x mean_x' sigmoid(f([x, x'])) -> discrete_space[winner].                     [dts.py:88-97]"""
    torch = nat.require_cuda()
    dev = f_vals.device if isinstance(f_vals, torch.Tensor) and f_vals.is_cuda else _common.device_of(device)
    f = _common.as_f32_device(f_vals, dev).reshape(-1)
    n = int(np.sqrt(f.shape[0]))
    score = torch.empty(n, dtype=torch.float32, device=dev)
    win = torch.empty(1, dtype=torch.int64, device=dev)
    nat.check(nat.lib().tkbo_soft_copeland(nat.handle(dev.index), nat.ptr(f), n, nat.ptr(score), nat.ptr(win),
                                           nat.current_stream_ptr(dev)), "tkbo_soft_copeland")
    winner = discrete_space[int(win.item())]
    return (winner, score.cpu()) if return_scores else winner


def thompson_copeland_winners(m, query_points, discrete_space, D=1000, count=1, *, device=None, return_scores=False):
    """``count`` independent rounds of sample_f + soft_copeland_maximizer (dts.py:71-97) in one fused launch: the
    ``count`` posterior samples share one (W, b) draw, the n^2 duels [x, x'] of ``discrete_space`` are generated inside
    the kernel and reduced to soft-Copeland scores on the fly (never stored).  Returns the (count, d) winners
    (and the (count, n) scores)."""
    pts = to_numpy(discrete_space)
    basis = fourier_features.SharedBasis.sample(to_numpy(query_points), m, D=D, count=count, device=device)
    score, win = basis.duel_copeland(pts, scores=return_scores)
    winners = pts[win.cpu().numpy()]
    return (winners, score.cpu()) if return_scores else winners
