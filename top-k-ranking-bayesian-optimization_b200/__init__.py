"""B200-native drop-in for the RFF maximiser-sampling + MPES hot path of
sebtsh/Top-k-Ranking-Bayesian-Optimization.

Import it the way the reference's scripts import their package (experiments/MPES/MPES_Forr.py:22-24):

    PBO = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    PBO.fourier_features.sample_maximizers(X=..., count=..., n_init=..., D=..., model=..., min_val=..., max_val=...)
    PBO.acquisitions.pes.I_batch(samples, maximizers, model)
    PBO.acquisitions.rank_pes.I_batch(chi, x_star, model, topk=3, num_samples=1024)
    PBO.acquisitions.dts.sample_f(model, X, combs, D)

Only the modules on that path exist (same sub-module layout as the reference's ``__init__.py:1``):
``fourier_features``, ``acquisitions.{pes,rank_pes,dts}``, plus ``model`` (a duck-typed stand-in for the
GPflow SVGP the reference reads ``kernel`` / ``q_mu`` / ``q_sqrt`` from) and ``distributed`` (candidate
sharding across GPUs).  All arithmetic on the path runs in hand-written sm_100a kernels from
``libtkbo.so`` (C ABI: include/tkbo.h) bound with ctypes; there is no TensorFlow, Triton or CPU fallback.
"""
from . import _native  # noqa: F401
from . import model, fourier_features, acquisitions, distributed  # noqa: F401
from ._native import TkboError  # noqa: F401

__all__ = ["fourier_features", "acquisitions", "model", "distributed", "TkboError"]
