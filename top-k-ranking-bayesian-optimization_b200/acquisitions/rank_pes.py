"""Multinomial PES for top-k ranking queries — signatures of the reference's acquisitions/rank_pes.py.

``I_batch`` keeps the reference call (rank_pes.py:10); the Monte-Carlo body (rank_pes.py:38-105) runs in the
K2 kernel ``tkbo_mpes_topk``.  The permutation helpers (rank_pes.py:180-242) are host-side integer code kept
for API parity; when all k-permutations are used the kernel enumerates them itself (the MI sums over every
observation type, so the table order is irrelevant) and only the >1000-permutation random-subset case
(rank_pes.py:49-53) ships an explicit table to the GPU.
"""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..model import to_numpy
from . import _common

N_MAX_PERMUTATION = 1000      # rank_pes.py:48


def I_batch(chi, x_star, model, topk=None, num_samples=10000, indifference_threshold=0.0, *, device=None,
            rff_features=None, f_samples=None):
    """Mutual information between each query's top-k ranking outcome and the maximiser.      [rank_pes.py:10-108]

    chi (num_data, num_choice, d), x_star (num_max, d) -> (num_data,) CPU float64 tensor.
    ``indifference_threshold`` is accepted and unused, as in the reference.  Extras: ``f_samples`` (S, Q*C + M)
    supplies pre-drawn joint samples in the reference's column order (rank_pes.py:25-35); ``rff_features`` forces
    RFF joint sampling with that many features instead of ``model.predict_f_samples``.
    Unlike rank_pes.py:40 (np.bincount without minlength, IndexError at :75) a trailing maximiser that wins no
    sample is simply dropped like any other empty group.
    """
    chi = to_numpy(chi)
    x_star = to_numpy(x_star)
    num_data, num_choice, d = chi.shape
    num_max = x_star.shape[0]
    if topk is None:
        topk = num_choice
    dev = _common.device_of(device)
    if f_samples is None:
        x_all = np.concatenate([chi.reshape(num_data * num_choice, d), x_star.reshape(num_max, d)], axis=0)   # :25
        fs = _common.joint_samples(model, x_all, num_samples, dev, rff_features)                              # :28
    else:
        fs = _common.as_f32_device(f_samples, dev)
    S = fs.shape[0]
    assert fs.shape[1] == num_data * num_choice + num_max, "f_samples must be (S, Q*C + M)"
    fstar = fs[:, num_data * num_choice:].contiguous()                                  # :31
    fchi = fs[:, :num_data * num_choice].contiguous().reshape(S, num_data, num_choice)  # :34-35

    n_permutations = precompute_n_permutations(num_choice, topk)
    n_total = int(n_permutations[topk])
    randomized = n_total > N_MAX_PERMUTATION                                            # :49
    perms = get_rand_permutation_k_in_n(num_choice, topk, N_MAX_PERMUTATION, n_permutations) if randomized else None
    print("Number of {} observations for top-{} in {} choices: {}".format(
        "randomized" if randomized else "all possible", topk, num_choice,
        N_MAX_PERMUTATION if randomized else n_total))                                  # :65-69
    mi = _common.mpes_from_device_samples(fchi, fstar, topk, nat.MPES_RANK, perms)
    return mi.cpu()


# ---- permutation helpers (host integers) --------------------------------------------------------
def precompute_n_permutations(n, k):
    """out[i] = (n-k+i)!/(n-k)!                                                          [rank_pes.py:202-215]"""
    assert k <= n
    out = np.ones(k + 1, dtype=int)
    for i in range(1, k + 1):
        out[i] = out[i - 1] * (n - k + i)
    return out


def get_ith_permutation_k_in_n(n, k, i, n_permutations):
    """Mixed-radix unranking of the i-th k-permutation of range(n), lexicographic.       [rank_pes.py:180-199]"""
    options = list(range(n))
    x = np.zeros(k, dtype=int)
    idx = int(i)
    for j in reversed(range(k)):
        opt_i = idx // int(n_permutations[j])
        x[k - 1 - j] = options.pop(opt_i)
        idx -= opt_i * int(n_permutations[j])
    return x


def get_all_permutation_k_in_n(n, k, n_permutations=None):
    """(n!/(n-k)!, k) table.                                                             [rank_pes.py:218-228]"""
    assert k <= n
    if n_permutations is None:
        n_permutations = precompute_n_permutations(n, k)
    return np.stack([get_ith_permutation_k_in_n(n, k, i, n_permutations) for i in range(int(n_permutations[k]))])


def get_rand_permutation_k_in_n(n, k, size, n_permutations=None):
    """`size` permutations drawn with replacement (np.random.randint, as the reference). [rank_pes.py:231-242]"""
    assert k <= n
    if n_permutations is None:
        n_permutations = precompute_n_permutations(n, k)
    randi = np.random.randint(low=0, high=n_permutations[k], size=size)
    return np.stack([get_ith_permutation_k_in_n(n, k, j, n_permutations) for j in randi])


def get_log_likelihood(fx, permutations, normalizer=None, *, device=None):
    """log (1/normalizer) sum_s P_s(order) for every row of ``permutations`` -> (P, Q) CPU float64 tensor.
    [rank_pes.py:111-134]  Provided for API parity; computed on the GPU with torch ops on (S, Q, C) samples
    (it is not on the fused path: I_batch never materialises this matrix)."""
    torch = nat.require_cuda()
    dev = _common.device_of(device)
    fx = torch.as_tensor(to_numpy(fx)).to(dev)
    S = fx.shape[0]
    normalizer = S if normalizer is None else normalizer
    out = []
    for order in np.asarray(permutations):
        out.append(get_log_likelihood_given_order(fx, order, normalizer, device=dev))
    return torch.stack(out).cpu()


def get_log_likelihood_given_order(fx, order, normalizer=None, *, device=None):
    """Plackett-Luce log-likelihood of one ascending-preference order -> (Q,).           [rank_pes.py:137-177]"""
    torch = nat.require_cuda()
    dev = _common.device_of(device)
    fx = torch.as_tensor(to_numpy(fx) if not isinstance(fx, torch.Tensor) else fx).to(dev).to(torch.float64)
    S, Q, C = fx.shape
    normalizer = S if normalizer is None else normalizer
    order = [int(o) for o in np.asarray(order)]
    pool = [c for c in range(C) if c not in set(order)]
    ll = torch.zeros((S, Q), dtype=torch.float64, device=dev)
    for o in order:
        pool = [o] + pool
        ll = ll + fx[:, :, o] - torch.logsumexp(fx[:, :, pool], dim=-1)
    return torch.logsumexp(ll - np.log(normalizer), dim=0)
