"""float64 NumPy restatement of the MPES / PES Monte-Carlo entropy estimates and the DTS tail.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Reference files:
``acquisitions/rank_pes.py`` (top-k ranking MPES), ``acquisitions/pes.py`` (top-1 PES/MPES),
``acquisitions/dts.py`` (soft-Copeland maximiser).  All estimators are written *given the
f-samples*: the reference draws them with ``model.predict_f_samples`` (GPflow) inside the call,
the deterministic core that follows is what is restated (and what the CUDA kernels replace).
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------------------
# rank_pes.py:180-228  k-permutation table (mixed-radix unranking, lexicographic order)
# --------------------------------------------------------------------------------------
def precompute_n_permutations(n, k):
    """out[i] = (n-k+i)! / (n-k)!  for i = 0..k          [rank_pes.py:202-215]"""
    assert k <= n
    out = np.ones(k + 1, dtype=np.int64)
    for i in range(1, k + 1):
        out[i] = out[i - 1] * (n - k + i)
    return out


def ith_permutation_k_in_n(n, k, i, n_perm):
    """rank_pes.py:180-199: digit j (most significant first) picks options[idx // n_perm[j]]."""
    options = list(range(n))
    out = np.zeros(k, dtype=np.int64)
    idx = int(i)
    for j in range(k - 1, -1, -1):
        opt = idx // int(n_perm[j])
        out[k - 1 - j] = options.pop(opt)
        idx -= opt * int(n_perm[j])
    return out


def all_permutations_k_in_n(n, k):
    """rank_pes.py:218-228.  (P, k) int table, P = n!/(n-k)!."""
    n_perm = precompute_n_permutations(n, k)
    return np.stack([ith_permutation_k_in_n(n, k, i, n_perm) for i in range(int(n_perm[k]))])


# --------------------------------------------------------------------------------------
# rank_pes.py:137-177  Plackett-Luce likelihood of one partial ranking
# --------------------------------------------------------------------------------------
def _logsumexp(a, axis):
    m = np.max(a, axis=axis, keepdims=True)
    return np.squeeze(m, axis=axis) + np.log(np.sum(np.exp(a - m), axis=axis))


def log_likelihood_given_order(fx, order, normalizer):
    """fx (S', Q, C); ``order`` ascending preference: order[i] is chosen from
    {unranked complement} + {order[0..i]} (rank_pes.py:160-166); then
    logsumexp over samples of (ll - log normalizer) (rank_pes.py:173).  -> (Q,)"""
    C = fx.shape[2]
    order = np.asarray(order, dtype=np.int64)
    pool = [c for c in range(C) if c not in set(order.tolist())]
    ll = np.zeros(fx.shape[:2])
    for i in range(order.shape[0]):
        pool = [int(order[i])] + pool
        ll += fx[:, :, order[i]] - _logsumexp(fx[:, :, pool], axis=-1)
    return _logsumexp(ll - np.log(normalizer), axis=0)


def log_likelihood(fx, perms, normalizer):
    """rank_pes.py:111-134 -> (P, Q)."""
    return np.stack([log_likelihood_given_order(fx, o, normalizer) for o in perms])


# --------------------------------------------------------------------------------------
# rank_pes.py:10-108  I_batch, given the samples
# --------------------------------------------------------------------------------------
def rank_mpes_from_samples(fchi, fstar, topk=None, perms=None):
    """Mutual information between the top-k ranking outcome of each query and the maximiser.

    fchi (S, Q, C), fstar (S, M) -> (Q,).  Mirrors rank_pes.py:38-105: group samples by
    argmax over the M maximisers, p(x*) = bincount / S, log p(z) and log p(z, x*) for every
    permutation (normaliser = S), drop maximisers with no sample, sum p(z,m)(log p(z|m) - log p(z)).
    ``perms`` overrides the table (the reference's >1000-permutation random subset case).
    """
    fchi = np.asarray(fchi, dtype=np.float64)
    fstar = np.asarray(fstar, dtype=np.float64)
    S, Q, C = fchi.shape
    M = fstar.shape[1]
    if topk is None:
        topk = C
    arg = np.argmax(fstar, axis=1)                                   # rank_pes.py:39
    counts = np.bincount(arg, minlength=M)
    p_xstar = counts / counts.sum()                                  # :40-41
    if perms is None:
        perms = all_permutations_k_in_n(C, topk)
    log_p_obs = log_likelihood(fchi, perms, S)                       # :58-60  (P, Q)
    mi = np.zeros(Q)
    for m in range(M):                                               # :74-82
        if p_xstar[m] <= 0:
            continue                                                 # :86-93 (dropped groups)
        log_p_joint = log_likelihood(fchi[arg == m], perms, S)       # (P, Q)
        log_p_cond = log_p_joint - np.log(p_xstar[m])                # :97
        mi += np.sum(np.exp(log_p_joint) * (log_p_cond - log_p_obs), axis=0)   # :101-105
    return mi


def rank_mpes_streaming(fchi, fstar_argmax, M, perms):
    """Closed-form / streaming formulation used by the CUDA kernel, in float64:
    with e = exp(f - max f), T = sum e, the Plackett-Luce probability of ``order`` is
    e[o_{k-1}]/T * e[o_{k-2}]/(T - e[o_{k-1}]) * ... ; accumulate per (group, permutation),
    then MI = sum_{m,z} p(z,m) log( p(z,m) / (p(m) p(z)) ).  Algebraically identical to
    ``rank_mpes_from_samples`` (empty groups contribute nothing)."""
    fchi = np.asarray(fchi, dtype=np.float64)
    S, Q, C = fchi.shape
    perms = np.asarray(perms, dtype=np.int64)
    P, k = perms.shape
    e = np.exp(fchi - fchi.max(axis=2, keepdims=True))
    T = e.sum(axis=2)
    prob = np.ones((P, S, Q))
    for p in range(P):
        rem = T.copy()
        for i in range(k - 1, -1, -1):
            c = perms[p, i]
            prob[p] *= e[:, :, c] / rem
            rem = rem - e[:, :, c]
    arg = np.asarray(fstar_argmax)
    p_z = prob.sum(axis=1) / S                                       # (P, Q)
    mi = np.zeros(Q)
    for m in range(M):
        sel = arg == m
        cnt = int(sel.sum())
        if cnt == 0:
            continue
        p_zm = prob[:, sel, :].sum(axis=1) / S
        mi += np.sum(p_zm * np.log(p_zm / ((cnt / S) * p_z)), axis=0)
    return mi


# --------------------------------------------------------------------------------------
# pes.py:44-123  top-1-of-C MI with the epsilon-smoothed counts, given the samples
# --------------------------------------------------------------------------------------
KERAS_EPSILON = 1e-7      # tf.keras.backend.epsilon()


def pes_I_from_samples(samples, num_max):
    """samples (S, M + C): columns [f(x*_1..M), f(chi_1..C)] (pes.py:105-107) -> scalar.

    log p(m)   = log((count_m + eps) / (S + M eps))                          pes.py:66-70
    lik_i(s)   = exp(f_i) / sum_c exp(f_c)   (un-stabilised softmax)          pes.py:52-53
    log p(z_i | m) = log((sum_{s: argmax=m} lik_i(s) + eps) / S) - log p(m)   pes.py:85-89
    I = sum_m p(m) sum_i p(z_i|m) (log p(z_i|m) - logsumexp_m(log p(m) + log p(z_i|m)))   :115-123
    """
    samples = np.asarray(samples, dtype=np.float64)
    S = samples.shape[0]
    M = int(num_max)
    C = samples.shape[1] - M
    arg = np.argmax(samples[:, :M], axis=1)
    count = np.zeros(M) + KERAS_EPSILON + np.bincount(arg, minlength=M)
    log_p_m = np.log(count / (S + M * KERAS_EPSILON))
    fz = samples[:, M:]
    total = 0.0
    expected = np.zeros((M, C))
    for i in range(C):
        lik = np.exp(fz[:, i]) / np.sum(np.exp(fz), axis=1)
        grouped = np.zeros(M)
        np.add.at(grouped, arg, lik)                                 # tf.scatter_nd sums duplicates
        grouped = grouped + KERAS_EPSILON
        log_p_z_m = np.log(grouped / S) - log_p_m
        norm = _logsumexp(log_p_m + log_p_z_m, axis=0)
        expected[:, i] = np.exp(log_p_z_m) * (log_p_z_m - norm)
    total = np.sum(np.exp(log_p_m) * expected.sum(axis=1))
    return float(total)


def pes_I_batch_from_samples(samples_batch, num_max):
    """pes.py:126-132 (one independent sample matrix per query, as the reference draws them)."""
    return np.array([pes_I_from_samples(s, num_max) for s in samples_batch])


# --------------------------------------------------------------------------------------
# dts.py:85-97
# --------------------------------------------------------------------------------------
def logistic(x):
    return 1.0 / (1.0 + np.exp(-x))                                  # dts.py:48-49


def soft_copeland_scores(f_vals):
    """mean over x' of sigma(f([x, x'])) for every x; f_vals (n^2, 1) i-major.   dts.py:94-96"""
    f_vals = np.asarray(f_vals, dtype=np.float64)
    n = int(np.sqrt(f_vals.shape[0]))
    return np.mean(np.reshape(logistic(f_vals), [n, n]), axis=1)


def soft_copeland_maximizer(f_vals, discrete_space):
    """dts.py:88-97."""
    return np.asarray(discrete_space)[np.argmax(soft_copeland_scores(f_vals))]


def dts_sample_f(W, b, query_points, q_mu, q_sqrt, z, combs):
    """dts.py:80-85 with the random draws (W, b, z) injected: theta is fitted on Phi(query_points)
    with the same basis, then f = Phi(combs) theta.  W (1, D, 2d), b (1, D, 1), z (1, n) -> (n^2, 1)."""
    from .rff_oracle import fourier_features, sample_theta_variational
    phi = fourier_features(np.asarray(combs)[None], W, b)
    phi_y = fourier_features(np.asarray(query_points)[None], W, b)
    theta = sample_theta_variational(phi_y, q_mu, q_sqrt, z=z)
    return np.squeeze(phi @ theta, axis=0)


# ---------------------------------------------------------------------------------------------------
# top-1 MPES with an indifference outcome (acquisitions/indiff_pes.py)
# ---------------------------------------------------------------------------------------------------
def indiff_log_likelihood(fx, normalizer=None, indifference_threshold=0.0):
    """(C + 1, Q): log (1/normalizer) sum_s P_s(z), z = "choice i is the most preferred" (i < C) or "none".
    fx (S, Q, C).                                                                       [indiff_pes.py:95-130]"""
    fx = np.asarray(fx, dtype=np.float64)
    S, Q, C = fx.shape
    normalizer = S if normalizer is None else normalizer
    indiff = (1.0 - np.eye(C)) * indifference_threshold                                 # :105-106
    base = fx[..., :, None] + indiff                                                    # :108, :111 (S, Q, C, C)
    ll_pref = fx - _logsumexp(base, axis=-2)                                            # :110-113
    p_ind = np.clip(1.0 - np.exp(ll_pref).sum(axis=-1, keepdims=True), 1e-100, 1.0 - 1e-100)   # :116-117
    ll = np.concatenate([ll_pref, np.log(p_ind)], axis=-1)                              # :121
    return _logsumexp(ll - np.log(normalizer), axis=0).T                                # :126-129


def indiff_mpes_from_samples(fchi, fstar, indifference_threshold=0.1):
    """MI (Q,) of indiff_pes.I_batch given the joint samples fchi (S, Q, C), fstar (S, M).  [indiff_pes.py:41-92]"""
    fchi = np.asarray(fchi, dtype=np.float64)
    fstar = np.asarray(fstar, dtype=np.float64)
    S, M = fstar.shape
    arg = np.argmax(fstar, axis=1)                                                      # :42
    count = np.bincount(arg, minlength=M)                                               # :43 (minlength: see rank_pes note)
    p_xstar = count / count.sum()                                                       # :44
    log_p_obs = indiff_log_likelihood(fchi, S, indifference_threshold)                  # :48-50
    mi = np.zeros(fchi.shape[1])
    for m in range(M):
        if p_xstar[m] <= 0:                                                             # :58, :68-76
            continue
        ll = indiff_log_likelihood(fchi[arg == m], S, indifference_threshold)           # :60-65 (C + 1, Q)
        mi += np.sum(np.exp(ll) * (ll - np.log(p_xstar[m]) - log_p_obs), axis=0)        # :80-90
    return mi


# --------------------------------------------------------------------------------------
# pes.py:9-41  grid-argmax maximiser samplers, given the per-batch draws
# --------------------------------------------------------------------------------------
def grid_maximizers_from_batches(data, batches, batch_size):
    """pes.py:24-41: ``samples`` (count, num_data) is filled batch by batch with independent joint draws
    (``batches[i]`` (count, batch_size) = what ``model.predict_f_samples(data[first:last], count)`` returned), rows past
    the last FULL batch keep their initial zeros (the loop runs ``num_data // batch_size`` times), then the per-sample
    argmax over all num_data columns picks the maximiser.  -> (count, input_dims)"""
    data = np.asarray(data, dtype=np.float64)
    num_data = data.shape[0]
    count = np.asarray(batches[0]).shape[0]
    samples = np.zeros((count, num_data))
    for i in range(num_data // batch_size):
        samples[:, i * batch_size:(i + 1) * batch_size] = np.asarray(batches[i], dtype=np.float64)
    return np.take(data, np.argmax(samples, axis=1), axis=0)
