"""Predictive entropy search for top-1 queries with an indifference outcome - signatures of the reference's
acquisitions/indiff_pes.py.  A query (x_0 .. x_{C-1}) has C + 1 possible observations: "x_i is the most preferred"
or "cannot tell" (indiff_pes.py:1-9).  ``I_batch`` keeps the reference call (indiff_pes.py:15); the Monte-Carlo body
(:44-92) and the likelihood (:95-130) run in the ``tkbo_mpes_indiff`` kernel.
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..model import to_numpy
from . import _common


def I_batch(chi, x_star, model, num_samples=1000, indifference_threshold=0.1, *, device=None, rff_features=None,
            f_samples=None):
    """Mutual information between each query's (top-1 or indifferent) outcome and the maximiser.  [indiff_pes.py:15-92]

    chi (num_data, num_choice, d), x_star (num_max, d) -> (num_data,) CPU float64 tensor.  Extras as in
    rank_pes.I_batch: ``f_samples`` (S, Q*C + M) pre-drawn joint samples in the reference's column order (:27-39),
    ``rff_features`` forces RFF joint sampling instead of ``model.predict_f_samples``."""
    torch = nat.require_cuda()
    chi = to_numpy(chi)
    x_star = to_numpy(x_star)
    num_data, num_choice, d = chi.shape
    num_max = x_star.shape[0]
    dev = _common.device_of(device)
    if f_samples is None:
        x_all = np.concatenate([chi.reshape(num_data * num_choice, d), x_star.reshape(num_max, d)], axis=0)   # :27
        fs = _common.joint_samples(model, x_all, num_samples, dev, rff_features)                              # :30
    else:
        fs = _common.as_f32_device(f_samples, dev)
    S = fs.shape[0]
    assert fs.shape[1] == num_data * num_choice + num_max, "f_samples must be (S, Q*C + M)"
    fstar = fs[:, num_data * num_choice:].contiguous()                                  # :33
    fchi = fs[:, :num_data * num_choice].contiguous()                                   # :36-37, (S, Q, C) row-major
    h = nat.handle(dev.index)
    st = nat.current_stream_ptr(dev)
    arg = torch.empty(S, dtype=torch.int32, device=dev)
    nat.check(nat.lib().tkbo_argmax_rows(h, nat.ptr(fstar), S, num_max, nat.ptr(arg), st), "tkbo_argmax_rows")   # :42
    mi = torch.empty(num_data, dtype=torch.float64, device=dev)
    nat.check(nat.lib().tkbo_mpes_indiff(h, nat.ptr(fchi), nat.ptr(arg), S, num_data, num_choice, num_max,
                                         float(indifference_threshold), nat.ptr(mi), st), "tkbo_mpes_indiff")
    return mi.cpu()


def get_log_likelihood(fx, normalizer=None, indifference_threshold=0.0, *, device=None):
    """log (1/normalizer) sum_s P_s(z) for the C + 1 observation types -> (C + 1, Q) CPU float64 tensor.
    [indiff_pes.py:95-130]  API parity only (I_batch never materialises it); torch ops on the GPU in float64."""
    torch = nat.require_cuda()
    dev = _common.device_of(device)
    fx = torch.as_tensor(to_numpy(fx) if not isinstance(fx, torch.Tensor) else fx).to(dev).to(torch.float64)
    S, Q, C = fx.shape
    normalizer = S if normalizer is None else normalizer
    indiff = (1.0 - torch.eye(C, dtype=torch.float64, device=dev)) * float(indifference_threshold)
    base = fx[..., :, None] + indiff                                   # (S, Q, C, C): row j, column i
    ll_pref = fx - torch.logsumexp(base, dim=-2)                        # :110-113
    p_ind = (1.0 - torch.exp(ll_pref).sum(dim=-1, keepdim=True)).clamp(1e-100, 1.0 - 1e-100)          # :116-117
    ll = torch.cat([ll_pref, torch.log(p_ind)], dim=-1)                 # (S, Q, C + 1)
    return torch.logsumexp(ll - np.log(normalizer), dim=0).T.cpu()      # :126-129
