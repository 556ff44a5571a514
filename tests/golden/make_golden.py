#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE SOURCES in /root/reference.

Run from the repo root in the build container:  python tests/golden/make_golden.py
(the GPU box has no /root/reference; it only reads the committed .npz files).

``rank_pes.py`` and ``indiff_pes.py`` run as shipped (NumPy/SciPy).  ``fourier_features.py``, ``pes.py`` and
``dts.py`` run under ``oracle/tf_shim.py`` (NumPy stand-in for the TF/TFP/GPflow symbols they call; TF 2.1 is not
installable offline; the gradient of the ascent loop ff.py:149-157 comes from torch autograd on the reference's own
loss closure).  Each file stores the exact inputs the reference function saw and the output
it returned, so both the oracle restatement (CPU tests) and the CUDA path (GPU tests) can be checked
against reference-produced numbers.
"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader, tf_shim  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def se_kernel(X, ls):
    d2 = ((X[:, None, :] - X[None, :, :]) / ls) ** 2
    return np.exp(-0.5 * d2.sum(-1))


def make_model(rng, n, d, lo, hi, ls):
    """Synthetic SVGP-like posterior: q_mu = L z, q_sqrt = 0.3 L, L = chol(K_SE + 1e-6 I)."""
    Xtr = rng.uniform(lo, hi, size=(n, d))
    L = np.linalg.cholesky(se_kernel(Xtr, ls) + 1e-6 * np.eye(n))
    q_mu = L @ rng.standard_normal((n, 1))
    q_sqrt = (0.3 * L)[None]
    return Xtr, q_mu, q_sqrt


def golden_fourier_features(ref):
    """ff.py:7-21, 56-87 on explicit inputs; both branches of ff.py:80-82."""
    rng = np.random.default_rng(101)
    out = {}
    for tag, (count, n, d, D) in {"under": (3, 7, 2, 48), "over": (2, 40, 3, 16)}.items():
        X = rng.uniform(-1.5, 1.5, size=(count, n, d))
        W = rng.standard_normal((count, D, d)) / 0.6
        b = rng.uniform(0, 2 * np.pi, size=(count, D, 1))
        phi = ref.fourier_features.fourier_features(tf_shim._t(X), tf_shim._t(W), tf_shim._t(b)).numpy()
        Xtr, q_mu, q_sqrt = make_model(rng, n, d, -1.5, 1.5, 0.6)
        tf_shim.seed(7)
        z = tf_shim._rng.standard_normal((count, n))          # the draw ff.py:71 will make next
        tf_shim.seed(7)
        theta = ref.fourier_features.sample_theta_variational(
            tf_shim._t(phi), tf_shim._t(q_mu), tf_shim._t(q_sqrt)).numpy()
        out.update({f"{tag}_X": X, f"{tag}_W": W, f"{tag}_b": b, f"{tag}_phi": phi,
                    f"{tag}_q_mu": q_mu, f"{tag}_q_sqrt": q_sqrt, f"{tag}_z": z, f"{tag}_theta": theta})
    np.savez_compressed(os.path.join(OUT, "ff_features_theta.npz"), **out)


def golden_sample_maximizers(ref):
    """ff.py:110-169 end to end with num_steps=0 (no autodiff in the shim): basis + theta sampling
    through sample_features_weights, x_star init, final eval / argmax / gather.  c1-like sizes
    (MPES_Forr.py:70-79: count=20, n_init=50, D=1000, d=1) and a 2-D case."""
    out = {}
    ffm = ref.fourier_features
    for tag, (count, n_init, D, n, d, lo, hi, ls) in {
            "forr": (20, 50, 1000, 12, 1, 0.0, 1.0, 0.2),
            "shc": (8, 64, 256, 20, 2, -1.5, 1.5, 0.6)}.items():
        rng = np.random.default_rng(202 + d)
        Xtr, q_mu, q_sqrt = make_model(rng, n, d, lo, hi, ls)
        model = tf_shim.StubModel(kernel=tf_shim.RBFKernel(np.full(d, ls)), q_mu=q_mu, q_sqrt=q_sqrt)
        captured = {}
        orig = ffm.sample_features_weights

        def spy(X, model, D, _orig=orig, _cap=captured):
            r = _orig(X, model, D)
            _cap["phi"], _cap["W"], _cap["b"], _cap["theta"] = [np.asarray(a) for a in r]
            return r

        ffm.sample_features_weights = spy
        tf_shim.seed(31 + d)
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                maxi = ffm.sample_maximizers(X=tf_shim._t(Xtr), count=count, n_init=n_init, D=D, model=model,
                                             min_val=lo, max_val=hi, num_steps=0).numpy()
        finally:
            ffm.sample_features_weights = orig
        x_star = np.asarray(tf_shim.created_variables()[-1])
        out.update({f"{tag}_Xtr": Xtr, f"{tag}_q_mu": q_mu, f"{tag}_q_sqrt": q_sqrt,
                    f"{tag}_W": captured["W"], f"{tag}_b": captured["b"], f"{tag}_theta": captured["theta"],
                    f"{tag}_x_star": x_star, f"{tag}_maximizers": maxi})
    np.savez_compressed(os.path.join(OUT, "ff_sample_maximizers.npz"), **out)


def golden_ascent(ref):
    """ff.py:110-169 with num_steps > 0: the reference's own optimisation loop (loss closure, optimizer.minimize,
    loss bookkeeping, early stop, final eval / argmax / gather) under the shim's autograd RMSprop.  Stores the basis,
    theta, the initial x_star, the final x_star, the maximizers and the printed step log."""
    out = {}
    ffm = ref.fourier_features
    cases = {  # tag: (count, n_init, D, n, d, lo, hi, ls, num_steps)
        "forr": (20, 50, 1000, 12, 1, 0.0, 1.0, 0.2, 40),          # c1 sizes (MPES_Forr.py:70-79), 40 steps
        "shc": (8, 64, 256, 20, 2, -1.5, 1.5, 0.6, 25),
        "stop": (3, 4, 32, 6, 1, 0.0, 0.0015, 0.2, 50)}            # every start is clipped after two steps -> early stop
    for tag, (count, n_init, D, n, d, lo, hi, ls, steps) in cases.items():
        rng = np.random.default_rng(303 + d + steps)
        Xtr, q_mu, q_sqrt = make_model(rng, n, d, lo, hi if hi > 1e-2 else 1.0, ls)
        model = tf_shim.StubModel(kernel=tf_shim.RBFKernel(np.full(d, ls)), q_mu=q_mu, q_sqrt=q_sqrt)
        captured = {}
        orig = ffm.sample_features_weights

        def spy(X, model, D, _orig=orig, _cap=captured):
            r = _orig(X, model, D)
            _cap["W"], _cap["b"], _cap["theta"] = [np.asarray(a) for a in r[1:]]
            return r

        ffm.sample_features_weights = spy
        tf_shim.seed(41 + d)
        log = io.StringIO()
        try:
            with contextlib.redirect_stdout(log):
                maxi = ffm.sample_maximizers(X=tf_shim._t(Xtr), count=count, n_init=n_init, D=D, model=model,
                                             min_val=lo, max_val=hi, num_steps=steps).numpy()
        finally:
            ffm.sample_features_weights = orig
        var = tf_shim.created_variables()[-1]
        # steps actually taken: the loop prints "Loss at step i" at i % 500 == 0 and again when it breaks (ff.py:152-156)
        lines = [l for l in log.getvalue().splitlines() if l.startswith("Loss at step")]
        last = int(lines[-1].split()[3].rstrip(":"))
        steps_run = steps if (last == 0 and len(lines) == 1) else last + 1
        out.update({f"{tag}_W": captured["W"], f"{tag}_b": captured["b"], f"{tag}_theta": captured["theta"],
                    f"{tag}_x0": var.initial_value, f"{tag}_x_final": np.asarray(var).copy(), f"{tag}_maximizers": maxi,
                    f"{tag}_meta": np.array([count, n_init, D, d, steps, steps_run]),
                    f"{tag}_bounds": np.array([lo, hi])})
    np.savez_compressed(os.path.join(OUT, "ff_ascent.npz"), **out)


def gp_like_samples(rng, S, n_points, d, amp, D=200, ls=0.3):
    """Smooth, GP-like joint samples at n_points uniform inputs: a fixed mean function plus amp * (RFF function draws).
    Small amp = a concentrated posterior (small q_sqrt): the winning maximiser says little about the query outcome and
    the MI lands in the 1e-4 .. 2e-2 range SURVEY 7 names as typical."""
    X = rng.uniform(0, 1, size=(n_points, d))
    W = rng.standard_normal((D, d)) / ls
    b = rng.uniform(0, 2 * np.pi, D)
    phi = np.sqrt(2.0 / D) * np.cos(X @ W.T + b)
    mean = phi @ rng.standard_normal(D)
    return mean[None, :] + amp * (rng.standard_normal((S, D)) @ phi.T)


def golden_rank_pes(ref):
    """rank_pes.I_batch as shipped (rank_pes.py:10-108) on fixed fp32-representable samples."""
    out = {}
    for tag, (S, Q, C, k, M) in {"c5k3": (512, 24, 5, 3, 20), "c4k4": (256, 10, 4, 4, 6),
                                 "c3k1": (300, 12, 3, 1, 5), "c5k2_sparse": (200, 8, 5, 2, 40)}.items():
        rng = np.random.default_rng({"c5k3": 1, "c4k4": 2, "c3k1": 3, "c5k2_sparse": 4}[tag])
        d = 6
        chi = rng.uniform(0, 1, size=(Q, C, d))
        x_star = rng.uniform(0, 1, size=(M, d))
        # correlated samples so that MI is non-trivial: low-rank function draws + noise
        basis = rng.standard_normal((Q * C + M, 6))
        fs = (rng.standard_normal((S, 6)) @ basis.T + 0.3 * rng.standard_normal((S, Q * C + M)))
        # rank_pes.py:40 calls np.bincount without minlength, so the reference raises IndexError at
        # rank_pes.py:75 whenever the LAST maximiser wins no sample; keep the fixtures inside the
        # domain where the reference runs by letting x*_{M-1} win sample 0 (empty groups stay interior).
        fs[0, Q * C + M - 1] = fs[0, Q * C:].max() + 1.0
        fs = fs.astype(np.float32).astype(np.float64)          # exactly what the fp32 GPU kernel consumes
        model = tf_shim.StubModel(fixed_samples=fs)
        with contextlib.redirect_stdout(io.StringIO()):
            mi = ref.rank_pes.I_batch(chi, x_star, model, topk=k, num_samples=S)
        out.update({f"{tag}_fsample_all": fs.astype(np.float32), f"{tag}_mi": mi,
                    f"{tag}_meta": np.array([S, Q, C, k, M])})
    # small-MI regime (GP-like samples) and a query whose best choice is > 100 above the rest (per-stage renormalisation,
    # rank_pes.py:166): the places where a float32 kernel is most likely to leave the 1e-4 relative tolerance
    for tag, (S, Q, C, k, M, amp, seed) in {"gp_small": (1024, 24, 5, 3, 20, 0.5, 21), "gp_tiny": (1024, 16, 5, 3, 20, 0.25, 22),
                                            "gp_c4k2": (768, 20, 4, 2, 12, 0.4, 23), "spread": (256, 12, 5, 3, 8, 0.5, 24)}.items():
        rng = np.random.default_rng(seed)
        d = 6
        chi = rng.uniform(0, 1, size=(Q, C, d))
        x_star = rng.uniform(0, 1, size=(M, d))
        fs = gp_like_samples(rng, S, Q * C + M, d, amp)
        if tag == "spread":
            fs[:, 0 * C + 2] += 120.0                 # query 0: choice 2 dominates by > 87 (float32 exp underflow)
            fs[:, 1 * C + 0] += 150.0                 # query 1: two tiers
            fs[:, 1 * C + 3] += 40.0
            fs[:, 2 * C + 4] -= 110.0                 # query 2: one choice far below
        fs[0, Q * C + M - 1] = fs[0, Q * C:].max() + 1.0
        fs = fs.astype(np.float32).astype(np.float64)
        model = tf_shim.StubModel(fixed_samples=fs)
        with contextlib.redirect_stdout(io.StringIO()):
            mi = ref.rank_pes.I_batch(chi, x_star, model, topk=k, num_samples=S)
        out.update({f"{tag}_fsample_all": fs.astype(np.float32), f"{tag}_mi": mi,
                    f"{tag}_meta": np.array([S, Q, C, k, M])})
    # get_log_likelihood (rank_pes.py:111-134) on its own: (P, Q) table of log p(z), normaliser = S
    fs = out["c5k3_fsample_all"].astype(np.float64)
    S, Q, C, k, M = [int(v) for v in out["c5k3_meta"]]
    perms = ref.rank_pes.get_all_permutation_k_in_n(C, k)
    out["c5k3_loglik"] = ref.rank_pes.get_log_likelihood(fs[:, :Q * C].reshape(S, Q, C), perms, normalizer=S)
    out["c5k3_loglik_order7"] = ref.rank_pes.get_log_likelihood_given_order(fs[:, :Q * C].reshape(S, Q, C), perms[7], normalizer=S)
    out["perms_5_3"] = ref.rank_pes.get_all_permutation_k_in_n(5, 3)
    out["nperm_5_3"] = ref.rank_pes.precompute_n_permutations(5, 3)
    out["perms_4_2"] = ref.rank_pes.get_all_permutation_k_in_n(4, 2)
    np.savez_compressed(os.path.join(OUT, "rank_pes.npz"), **out)


def golden_indiff_pes(ref):
    """indiff_pes.I_batch as shipped (indiff_pes.py:15-92) on fixed fp32-representable samples."""
    out = {}
    for tag, (S, Q, C, M, thr) in {"c3": (400, 16, 3, 8, 0.1), "c5": (256, 10, 5, 12, 0.5), "c2_t0": (300, 12, 2, 6, 0.0)}.items():
        rng = np.random.default_rng({"c3": 11, "c5": 12, "c2_t0": 13}[tag])
        d = 4
        chi = rng.uniform(0, 1, size=(Q, C, d))
        x_star = rng.uniform(0, 1, size=(M, d))
        basis = rng.standard_normal((Q * C + M, 5))
        fs = (rng.standard_normal((S, 5)) @ basis.T + 0.3 * rng.standard_normal((S, Q * C + M)))
        # indiff_pes.py:43 also calls np.bincount without minlength: let the last maximiser win sample 0
        fs[0, Q * C + M - 1] = fs[0, Q * C:].max() + 1.0
        fs = fs.astype(np.float32).astype(np.float64)
        model = tf_shim.StubModel(fixed_samples=fs)
        mi = ref.indiff_pes.I_batch(chi, x_star, model, num_samples=S, indifference_threshold=thr)
        ll = ref.indiff_pes.get_log_likelihood(fs[:, :Q * C].reshape(S, Q, C), normalizer=S, indifference_threshold=thr)
        out.update({f"{tag}_fsample_all": fs.astype(np.float32), f"{tag}_mi": mi, f"{tag}_loglik": ll,
                    f"{tag}_meta": np.array([S, Q, C, M]), f"{tag}_threshold": np.array(thr)})
    for tag, (S, Q, C, M, thr, amp, seed) in {"gp_small": (1024, 20, 3, 12, 0.1, 0.4, 31), "gp_tiny": (1000, 16, 2, 20, 0.05, 0.2, 32)}.items():
        rng = np.random.default_rng(seed)
        d = 4
        chi = rng.uniform(0, 1, size=(Q, C, d))
        x_star = rng.uniform(0, 1, size=(M, d))
        fs = gp_like_samples(rng, S, Q * C + M, d, amp)
        fs[0, Q * C + M - 1] = fs[0, Q * C:].max() + 1.0
        fs = fs.astype(np.float32).astype(np.float64)
        model = tf_shim.StubModel(fixed_samples=fs)
        mi = ref.indiff_pes.I_batch(chi, x_star, model, num_samples=S, indifference_threshold=thr)
        out.update({f"{tag}_fsample_all": fs.astype(np.float32), f"{tag}_mi": mi,
                    f"{tag}_meta": np.array([S, Q, C, M]), f"{tag}_threshold": np.array(thr)})
    np.savez_compressed(os.path.join(OUT, "indiff_pes.npz"), **out)


def golden_pes(ref):
    """pes.I / pes.I_batch (pes.py:92-132) with one fixed sample matrix per query. c1 sizes:
    S=1000, M=20, C=2 (MPES_Forr.py:70-79) plus a C=3 case."""
    out = {}
    for tag, (S, M, C, Qn) in {"c1": (1000, 20, 2, 6), "c3": (400, 7, 3, 4)}.items():
        rng = np.random.default_rng({"c1": 11, "c3": 12}[tag])
        basis = rng.standard_normal((Qn, M + C, 5))
        per_query = (rng.standard_normal((Qn, S, 5)) @ np.swapaxes(basis, 1, 2)
                     + 0.2 * rng.standard_normal((Qn, S, M + C)))
        per_query = per_query.astype(np.float32).astype(np.float64)
        model = tf_shim.StubModel(fixed_samples=lambda x, n, call, pq=per_query: pq[call])
        chi_batch = rng.uniform(0, 1, size=(Qn, C, 1))
        x_star = rng.uniform(0, 1, size=(M, 1))
        if S != 1000:
            vals = np.array([float(ref.pes.I(tf_shim._t(c), tf_shim._t(x_star), model, num_samples=S))
                             for c in chi_batch])
        else:
            vals = ref.pes.I_batch(tf_shim._t(chi_batch), tf_shim._t(x_star), model)
        out.update({f"{tag}_samples": per_query.astype(np.float32), f"{tag}_I": np.asarray(vals, dtype=np.float64),
                    f"{tag}_meta": np.array([S, M, C, Qn])})
    # small-MI regime: GP-like samples at [x*; chi] (pes.py:105 column order), c1 sizes
    S, M, C, Qn = 1000, 20, 2, 8
    rng = np.random.default_rng(41)
    per_query = np.stack([gp_like_samples(rng, S, M + C, 1, amp) for amp in (0.5, 0.4, 0.3, 0.25, 0.2, 0.5, 0.35, 0.3)])
    per_query = per_query.astype(np.float32).astype(np.float64)
    model = tf_shim.StubModel(fixed_samples=lambda x, n, call, pq=per_query: pq[call])
    chi_batch = rng.uniform(0, 1, size=(Qn, C, 1))
    x_star = rng.uniform(0, 1, size=(M, 1))
    vals = ref.pes.I_batch(tf_shim._t(chi_batch), tf_shim._t(x_star), model)
    out.update({"small_samples": per_query.astype(np.float32), "small_I": np.asarray(vals, dtype=np.float64),
                "small_meta": np.array([S, M, C, Qn])})
    np.savez_compressed(os.path.join(OUT, "pes.npz"), **out)


def golden_pes_maximizers(ref):
    """pes.sample_maximizers_discrete / _simple (pes.py:9-41) with a stub model returning fixed exact-GP-style batches:
    independent draws per 1000-point batch, rows beyond the last full batch stay zero (pes.py:33-38)."""
    rng = np.random.default_rng(51)
    count, num_data, d, batch = 12, 2300, 2, 1000
    data = rng.uniform(0, 1, size=(num_data, d))
    batches = [gp_like_samples(rng, count, batch, d, 0.6).astype(np.float32).astype(np.float64) - 0.5 for _ in range(num_data // batch)]
    model = tf_shim.StubModel(fixed_samples=lambda x, n, call, bs=batches: bs[call])
    maxi = np.asarray(ref.pes.sample_maximizers_discrete(model, count, data, batch_size=batch))
    # a second case in which every drawn value is negative: the zero tail rows win (reference behaviour, kept)
    neg = [b - 10.0 for b in batches]
    model2 = tf_shim.StubModel(fixed_samples=lambda x, n, call, bs=neg: bs[call])
    maxi_neg = np.asarray(ref.pes.sample_maximizers_discrete(model2, count, data, batch_size=batch))
    n_disc = 400
    simple = gp_like_samples(rng, count, n_disc, 1, 0.6).astype(np.float32).astype(np.float64)
    model3 = tf_shim.StubModel(fixed_samples=simple)
    maxi_simple = np.asarray(ref.pes.sample_maximizers_simple(model3, count, n_disc))
    np.savez_compressed(os.path.join(OUT, "pes_maximizers.npz"), data=data, batches=np.stack(batches).astype(np.float32),
                        maximizers=maxi, maximizers_neg=maxi_neg, simple_samples=simple.astype(np.float32),
                        maximizers_simple=maxi_simple, meta=np.array([count, num_data, d, batch, n_disc]))


def golden_pes_inputs(ref):
    """pes.sample_inputs / sample_inputs_discrete (pes.py:135-189) under a fixed np.random seed."""
    rng = np.random.default_rng(31)
    cur = rng.uniform(0, 1, size=(4, 3))
    data = rng.uniform(0, 1, size=(25, 3))
    np.random.seed(1234)
    a = np.asarray(ref.pes.sample_inputs(cur, 6, 3, -1.0, 2.0))
    np.random.seed(4321)
    b = np.asarray(ref.pes.sample_inputs_discrete(cur, data, 5, 4))
    np.savez_compressed(os.path.join(OUT, "pes_inputs.npz"), cur=cur, data=data, uniform=a, discrete=b)


def golden_dts(ref):
    """dts.sample_f (dts.py:71-85) with a Product kernel, and soft_copeland_maximizer (dts.py:88-97)."""
    import gpflow
    rng = np.random.default_rng(77)
    d_obj, n_disc, D, n = 2, 12, 64, 9
    grid = ref.dts.uniform_grid(d_obj, n_disc, 0, 1)                      # (144, 2)
    combs = ref.dts.combinations(grid)                                    # (144^2, 4)
    ls = np.array([0.25, 0.4])
    kernel = gpflow.kernels.Product([tf_shim.RBFKernel(ls[0]), tf_shim.RBFKernel(ls[1])])
    kernel.__class__ = gpflow.kernels.base.Product
    Xq, q_mu, q_sqrt = make_model(rng, n, 2 * d_obj, 0, 1, 0.4)
    model = tf_shim.StubModel(kernel=kernel, q_mu=q_mu, q_sqrt=q_sqrt)
    ffm = ref.fourier_features
    captured = {}
    orig = ffm.sample_fourier_features

    def spy(X, kernel, D=100):
        r = orig(X, kernel, D)
        captured["W"], captured["b"] = np.asarray(r[1]), np.asarray(r[2])
        return r

    ffm.sample_fourier_features = spy
    tf_shim.seed(5)
    try:
        f_vals = np.asarray(ref.dts.sample_f(model, Xq, combs, D))
    finally:
        ffm.sample_fourier_features = orig
    # the z drawn inside sample_theta_variational: replay the stream (W: D*2d normals, b: D uniforms, then z)
    tf_shim.seed(5)
    tf_shim._rng.standard_normal((1, D, 2 * d_obj))
    tf_shim._rng.uniform(0, 1, size=(1, D, 1))
    z = tf_shim._rng.standard_normal((1, n))
    winner = ref.dts.soft_copeland_maximizer(tf_shim._t(f_vals), grid)
    np.savez_compressed(os.path.join(OUT, "dts.npz"), grid=grid, combs=combs, lengthscales=ls, Xq=Xq,
                        q_mu=q_mu, q_sqrt=q_sqrt, W=captured["W"], b=captured["b"], z=z,
                        f_vals=f_vals, winner=np.asarray(winner))


def main():
    ref = ref_loader.load()
    golden_fourier_features(ref)
    golden_sample_maximizers(ref)
    golden_ascent(ref)
    golden_rank_pes(ref)
    golden_indiff_pes(ref)
    golden_pes(ref)
    golden_pes_maximizers(ref)
    golden_pes_inputs(ref)
    golden_dts(ref)
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
