"""Predictive Entropy Search (top-1 of C choices) — signatures of the reference's acquisitions/pes.py.

``I`` (pes.py:92-123) draws one sample matrix at [x_star; chi] and evaluates the eps-smoothed MI; the
Monte-Carlo body (pes.py:44-89, 108-123) runs in ``tkbo_mpes_topk`` (mode TKBO_MPES_PES).  ``I_batch``
(pes.py:126-132) is a Python loop over ``I`` in the reference, i.e. one independent sample matrix per query;
that is the default here too.  ``joint=True`` (opt-in; needs a model with ``predict_f_samples_device`` or
``rff_features``) lets all queries share one joint RFF draw at [x_star; chi_1; ...; chi_Q] scored by a single kernel
launch — the same estimator without 3000 tiny launches per BO iteration, but not the reference's sampling scheme.
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..model import to_numpy
from . import _common


def I(chi, x_star, model, num_samples=1000, *, device=None, f_samples=None):
    """chi (num_choices, d), x_star (num_max, d) -> scalar (0-d CPU float64 tensor).        [pes.py:92-123]
    ``f_samples`` (S, num_max + num_choices) supplies the matrix pes.py:106 would draw."""
    chi = to_numpy(chi)
    x_star = to_numpy(x_star)
    num_choices, num_max = chi.shape[0], x_star.shape[0]
    dev = _common.device_of(device)
    if f_samples is None:
        x_vals = np.concatenate([x_star, chi], axis=0)                       # pes.py:105
        fs = _common.joint_samples(model, x_vals, num_samples, dev)          # pes.py:106
    else:
        fs = _common.as_f32_device(f_samples, dev)
    S = fs.shape[0]
    fstar = fs[:, :num_max].contiguous()                                     # pes.py:107
    fchi = fs[:, num_max:].contiguous().reshape(S, 1, num_choices)
    return _common.mpes_from_device_samples(fchi, fstar, 1, nat.MPES_PES)[0].cpu()


def I_batch(chi_batch, x_star, model, *, num_samples=1000, joint=None, device=None, rff_features=None):
    """chi_batch (num_queries, num_choices, d) -> np.ndarray (num_queries,), as pes.py:126-132."""
    chi_batch = to_numpy(chi_batch)
    x_star = to_numpy(x_star)
    Q, C, d = chi_batch.shape
    M = x_star.shape[0]
    if joint is None:
        joint = rff_features is not None               # reference semantics unless the caller asks for the joint draw
    if not joint:
        return np.array([float(I(chi, x_star, model, num_samples=num_samples, device=device)) for chi in chi_batch])
    dev = _common.device_of(device)
    x_all = np.concatenate([x_star, chi_batch.reshape(Q * C, d)], axis=0)
    fs = _common.joint_samples(model, x_all, num_samples, dev, rff_features if rff_features is not None else 1024)
    S = fs.shape[0]
    fstar = fs[:, :M].contiguous()
    fchi = fs[:, M:].contiguous().reshape(S, Q, C)
    return _common.mpes_from_device_samples(fchi, fstar, 1, nat.MPES_PES).cpu().numpy()


def _argmax_rows_device(samples, dev):
    """Per-row arg max (lowest index on ties, np.argmax semantics) of a (count, num_data) matrix on the GPU."""
    torch = nat.require_cuda()
    smp = _common.as_f32_device(samples, dev)
    count, num_data = smp.shape
    arg = torch.empty(count, dtype=torch.int32, device=dev)
    nat.check(nat.lib().tkbo_argmax_rows(nat.handle(dev.index), nat.ptr(smp), count, num_data, nat.ptr(arg),
                                         nat.current_stream_ptr(dev)), "tkbo_argmax_rows")
    return arg.cpu().numpy().astype(np.int64)


def sample_maximizers_discrete(model, count, data, batch_size=1000, *, rff_features=None, device=None, return_info=False):
    """Grid-argmax maximiser sampling.                                                              [pes.py:24-41]

    Default = the reference's scheme: ``model.predict_f_samples(data[first:last], count)`` for every FULL batch of
    ``batch_size`` points (independent joint draws per batch; rows beyond the last full batch keep the value 0, as
    ``range(num_data // batch_size)`` leaves them), then the per-sample arg max over all ``num_data`` columns
    (``tkbo_argmax_rows``) and ``np.take`` -> (count, input_dims) CPU float64 tensor.

    ``rff_features=D`` (opt-in) instead draws ``count`` RFF posterior samples on ONE shared basis and evaluates them on
    all of ``data`` with the fused argmax kernel (``batch_size`` is then unused): a different sampler of the same
    posterior that scales to millions of points; ``return_info`` also returns the basis and the indices."""
    import torch
    data_np = to_numpy(data)
    dev = _common.device_of(device)
    if rff_features is not None:
        from .. import fourier_features as ff
        Z = model.Z if hasattr(model, "Z") else to_numpy(model.inducing_variable.Z)
        basis = ff.SharedBasis.sample(Z, model, D=int(rff_features), count=count, device=dev)
        val, idx = basis.argmax(data_np.astype(np.float32))
        out = torch.from_numpy(np.take(data_np, idx.cpu().numpy(), axis=0))
        return (out, {"basis": basis, "indices": idx.cpu(), "values": val.cpu()}) if return_info else out
    num_data = data_np.shape[0]
    samples = torch.zeros((count, num_data), dtype=torch.float32, device=dev)            # pes.py:33
    for i in range(num_data // batch_size):                                             # pes.py:35-38
        first, last = i * batch_size, (i + 1) * batch_size
        fs = model.predict_f_samples(data_np[first:last], count)
        samples[:, first:last] = _common.as_f32_device(fs, dev).reshape(count, batch_size)
    samples_argmax = _argmax_rows_device(samples, dev)                                  # pes.py:40
    out = torch.from_numpy(np.take(data_np, samples_argmax, axis=0))                    # pes.py:41
    return (out, {"indices": torch.from_numpy(samples_argmax)}) if return_info else out


def sample_maximizers_simple(model, count, num_discrete_points, *, rff_features=None, device=None):
    """Grid-argmax maximiser sampling on ``num_discrete_points`` points of [0, 1] -> (count, 1).     [pes.py:9-21]
    One joint draw ``model.predict_f_samples(xx, count)`` at all grid points (no batching in the reference)."""
    import torch
    xx = np.linspace(0.0, 1.0, num_discrete_points).reshape(num_discrete_points, 1)
    if rff_features is not None:
        return sample_maximizers_discrete(model, count, xx, rff_features=rff_features, device=device)
    dev = _common.device_of(device)
    samples = _common.as_f32_device(model.predict_f_samples(xx, count), dev).reshape(count, num_discrete_points)
    return torch.from_numpy(np.take(xx, _argmax_rows_device(samples, dev))[:, None])     # pes.py:20-21


# ---- the building blocks of I (pes.py:44-89): kept for API parity, plain torch float64 on the tensors' device; the
# ---- fused path (tkbo_mpes_topk, mode TKBO_MPES_PES) computes the same quantities without materialising them
EPSILON = 1e-7                 # tf.keras.backend.epsilon()


def samples_likelihood(z, f_z):
    """P(choice x is the most preferred given f) per sample, un-stabilised softmax as in the reference. [pes.py:45-53]"""
    import torch
    f_z = torch.as_tensor(f_z, dtype=torch.float64) if not isinstance(f_z, torch.Tensor) else f_z.to(torch.float64)
    x = int(z[0])
    return torch.exp(f_z[:, x]) / torch.exp(f_z).sum(dim=1)


def log_p_x_star_cond_D(num_max, samples_argmax):
    """log of the eps-smoothed fraction of samples each maximiser wins -> (num_max,).                 [pes.py:57-70]"""
    import torch
    arg = torch.as_tensor(samples_argmax).to(torch.int64)
    count = torch.bincount(arg, minlength=int(num_max)).to(torch.float64) + EPSILON
    return torch.log(count / (arg.shape[0] + num_max * EPSILON))


def log_p_z_cond_D_x_star(num_max, samples, samples_x_star_argmax, z, log_p_x_star_cond):
    """log p(z | D, x*) for every maximiser -> (num_max,).                                             [pes.py:74-89]"""
    import torch
    samples = torch.as_tensor(samples).to(torch.float64)
    arg = torch.as_tensor(samples_x_star_argmax).to(torch.int64).to(samples.device)
    lik = samples_likelihood(z, samples[:, num_max:])
    p = torch.zeros(int(num_max), dtype=torch.float64, device=samples.device).index_add_(0, arg, lik) + EPSILON
    return torch.log(p / samples.shape[0]) - torch.as_tensor(log_p_x_star_cond).to(samples.device)


def sample_inputs(current_inputs, num_samples, num_choices, min_val, max_val):
    """Query sets for the next observation: every existing input paired with the same ``num_samples`` uniform draws
    of the other ``num_choices - 1`` choices -> (num_samples * num_inputs, num_choices, input_dims) CPU float64
    tensor, row i * num_samples + j = [current_inputs[i], draw j].  Same ``np.random.uniform`` call as the
    reference, so a seeded run reproduces its query sets bit for bit.                       [pes.py:135-161]"""
    import torch                                           # host-side bookkeeping: no GPU involved
    cur = to_numpy(current_inputs)
    num_inputs, input_dims = cur.shape
    uniform_samples = np.random.uniform(low=min_val, high=max_val, size=(num_samples, num_choices - 1, input_dims))
    samples = np.empty((num_inputs, num_samples, num_choices, input_dims))
    samples[:, :, 0, :] = cur[:, None, :]
    samples[:, :, 1:, :] = uniform_samples[None]
    return torch.from_numpy(samples.reshape(num_inputs * num_samples, num_choices, input_dims))


def sample_inputs_discrete(current_inputs, data, num_samples, num_choices):
    """As sample_inputs with the other choices drawn (with replacement) from the rows of ``data``.   [pes.py:164-189]"""
    import torch
    cur = to_numpy(current_inputs)
    data = to_numpy(data)
    num_inputs, input_dims = cur.shape
    random_indices = np.random.choice(data.shape[0], (num_samples, num_choices - 1))
    data_samples = np.take(data, random_indices, axis=0)
    samples = np.empty((num_inputs, num_samples, num_choices, input_dims))
    samples[:, :, 0, :] = cur[:, None, :]
    samples[:, :, 1:, :] = data_samples[None]
    return torch.from_numpy(samples.reshape(num_inputs * num_samples, num_choices, input_dims))
