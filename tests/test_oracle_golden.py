"""Pin the oracle restatement against reference-produced golden vectors (CPU, no GPU).

The .npz files were produced by tests/golden/make_golden.py executing the unmodified reference
sources; when /root/reference is mounted the same comparison is also made live."""
import contextlib
import io

import numpy as np
import pytest

from oracle import mpes_oracle, ref_loader, rff_oracle, tf_shim


@pytest.mark.parametrize("tag", ["under", "over"])
def test_fourier_features_and_theta(golden, tag):
    g = golden("ff_features_theta.npz")
    phi = rff_oracle.fourier_features(g[f"{tag}_X"], g[f"{tag}_W"], g[f"{tag}_b"])
    np.testing.assert_allclose(phi, g[f"{tag}_phi"], rtol=0, atol=1e-15)
    theta = rff_oracle.sample_theta_variational(g[f"{tag}_phi"], g[f"{tag}_q_mu"], g[f"{tag}_q_sqrt"], z=g[f"{tag}_z"])
    scale = np.abs(g[f"{tag}_theta"]).max()
    np.testing.assert_allclose(theta, g[f"{tag}_theta"], rtol=0, atol=1e-9 * scale)


@pytest.mark.parametrize("tag", ["forr", "shc"])
def test_eval_argmax_gather(golden, tag):
    g = golden("ff_sample_maximizers.npz")
    fvals, idx, maxi = rff_oracle.eval_argmax_gather(g[f"{tag}_x_star"], g[f"{tag}_W"], g[f"{tag}_b"], g[f"{tag}_theta"])
    np.testing.assert_array_equal(maxi, g[f"{tag}_maximizers"])
    assert maxi.shape == (g[f"{tag}_x_star"].shape[0], g[f"{tag}_x_star"].shape[2])
    # theta reproduces from the stored basis with the oracle's own inversion (interpolation property)
    Xtr = g[f"{tag}_Xtr"]
    count = g[f"{tag}_W"].shape[0]
    phi_tr = rff_oracle.fourier_features(np.tile(Xtr[None], (count, 1, 1)), g[f"{tag}_W"], g[f"{tag}_b"])
    q = phi_tr @ g[f"{tag}_theta"]                       # min-norm interpolant hits the q-samples
    assert np.all(np.isfinite(q))


RANK_TAGS = ["c5k3", "c4k4", "c3k1", "c5k2_sparse", "gp_small", "gp_tiny", "gp_c4k2", "spread"]


@pytest.mark.parametrize("tag", RANK_TAGS)
def test_rank_mpes(golden, tag):
    g = golden("rank_pes.npz")
    S, Q, C, k, M = [int(v) for v in g[f"{tag}_meta"]]
    fs = g[f"{tag}_fsample_all"].astype(np.float64)
    fchi = fs[:, :Q * C].reshape(S, Q, C)
    fstar = fs[:, Q * C:]
    ref = g[f"{tag}_mi"]
    mi = mpes_oracle.rank_mpes_from_samples(fchi, fstar, topk=k)
    np.testing.assert_allclose(mi, ref, rtol=1e-9, atol=1e-13)
    perms = mpes_oracle.all_permutations_k_in_n(C, k)
    if tag != "spread":        # the closed form with one global max underflows there; the kernel recentres (mpes.cu kExpBias)
        mi2 = mpes_oracle.rank_mpes_streaming(fchi, np.argmax(fstar, axis=1), M, perms)
        np.testing.assert_allclose(mi2, ref, rtol=1e-8, atol=1e-12)
    assert np.all(ref > -1e-12)                          # MI >= 0


def test_rank_log_likelihood_table(golden):
    """get_log_likelihood / _given_order (rank_pes.py:111-177) on their own, against the reference's output."""
    g = golden("rank_pes.npz")
    S, Q, C, k, M = [int(v) for v in g["c5k3_meta"]]
    fchi = g["c5k3_fsample_all"].astype(np.float64)[:, :Q * C].reshape(S, Q, C)
    perms = g["perms_5_3"]
    np.testing.assert_allclose(mpes_oracle.log_likelihood(fchi, perms, S), g["c5k3_loglik"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(mpes_oracle.log_likelihood_given_order(fchi, perms[7], S), g["c5k3_loglik_order7"],
                               rtol=1e-12, atol=1e-13)


@pytest.mark.parametrize("tag", ["c3", "c5", "c2_t0", "gp_small", "gp_tiny"])
def test_indiff_mpes(golden, tag):
    """Restatement of indiff_pes.I_batch / get_log_likelihood against the reference's own output."""
    g = golden("indiff_pes.npz")
    S, Q, C, M = (int(v) for v in g[f"{tag}_meta"])
    thr = float(g[f"{tag}_threshold"])
    fs = g[f"{tag}_fsample_all"].astype(np.float64)
    fchi, fstar = fs[:, :Q * C].reshape(S, Q, C), fs[:, Q * C:]
    if f"{tag}_loglik" in g.files:
        ll, ll_ref = mpes_oracle.indiff_log_likelihood(fchi, S, thr), g[f"{tag}_loglik"]
        rows = C + 1 if thr > 0 else C        # threshold 0: P(none) = 1 - sum_i P(i) is pure rounding noise (~1e-16)
        np.testing.assert_allclose(ll[:rows], ll_ref[:rows], rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(mpes_oracle.indiff_mpes_from_samples(fchi, fstar, thr), g[f"{tag}_mi"], rtol=1e-9, atol=1e-12)


def test_permutation_tables(golden):
    g = golden("rank_pes.npz")
    np.testing.assert_array_equal(mpes_oracle.all_permutations_k_in_n(5, 3), g["perms_5_3"])
    np.testing.assert_array_equal(mpes_oracle.all_permutations_k_in_n(4, 2), g["perms_4_2"])
    np.testing.assert_array_equal(mpes_oracle.precompute_n_permutations(5, 3), g["nperm_5_3"])
    # hand-checkable KATs
    p = mpes_oracle.all_permutations_k_in_n(5, 3)
    assert p[:3].tolist() == [[0, 1, 2], [0, 1, 3], [0, 1, 4]]
    assert mpes_oracle.precompute_n_permutations(5, 3).tolist() == [1, 3, 12, 60]
    assert len({tuple(r) for r in p.tolist()}) == 60


@pytest.mark.parametrize("tag", ["c1", "c3", "small"])
def test_pes_I(golden, tag):
    g = golden("pes.npz")
    S, M, C, Qn = [int(v) for v in g[f"{tag}_meta"]]
    vals = mpes_oracle.pes_I_batch_from_samples(g[f"{tag}_samples"].astype(np.float64), M)
    np.testing.assert_allclose(vals, g[f"{tag}_I"], rtol=1e-9, atol=1e-13)


def test_pes_equals_rank_top1_without_eps(golden):
    """pes.I and rank_pes(topk=1) estimate the same MI; they differ only by the eps smoothing."""
    g = golden("pes.npz")
    S, M, C, Qn = [int(v) for v in g["c1_meta"]]
    smp = g["c1_samples"].astype(np.float64)
    for q in range(Qn):
        a = mpes_oracle.pes_I_from_samples(smp[q], M)
        b = mpes_oracle.rank_mpes_from_samples(smp[q][:, M:].reshape(S, 1, C), smp[q][:, :M], topk=1)[0]
        assert abs(a - b) < 1e-5


def test_dts(golden):
    g = golden("dts.npz")
    f = mpes_oracle.dts_sample_f(g["W"], g["b"], g["Xq"], g["q_mu"], g["q_sqrt"], g["z"], g["combs"])
    np.testing.assert_allclose(f, g["f_vals"], rtol=0, atol=1e-9 * np.abs(g["f_vals"]).max())
    np.testing.assert_array_equal(mpes_oracle.soft_copeland_maximizer(g["f_vals"], g["grid"]), g["winner"])
    ls = rff_oracle.product_kernel_lengthscales(g["lengthscales"], 4)
    assert ls.tolist() == [0.25, 0.4, 0.25, 0.4]
    # the basis the reference drew has the duplicated per-dim scale: column std ~ 1/ls
    assert g["W"].shape == (1, 64, 4)


def test_shared_basis_matches_per_sample_form():
    """Shared-basis helpers equal the reference-shaped (count, ...) functions when every sample
    shares one (W, b): column s of the shared evaluation == ff.py:162 for sample s."""
    rng = np.random.default_rng(0)
    n, d, D, S, N = 9, 2, 40, 5, 300
    Xtr = rng.uniform(-1, 1, (n, d))
    W = rng.standard_normal((D, d))
    b = rng.uniform(0, 2 * np.pi, D)
    L = np.linalg.cholesky(np.exp(-0.5 * ((Xtr[:, None] - Xtr[None]) ** 2).sum(-1)) + 1e-6 * np.eye(n))
    q_mu, q_sqrt = L @ rng.standard_normal((n, 1)), (0.3 * L)[None]
    Z = rng.standard_normal((S, n))
    phi_tr = rff_oracle.fourier_features(Xtr[None], W[None], b[None, :, None])[0]
    Theta = rff_oracle.theta_shared_basis(phi_tr, q_mu, q_sqrt, Z)
    Wc, bc = np.tile(W[None], (S, 1, 1)), np.tile(b[None, :, None], (S, 1, 1))
    theta_ref = rff_oracle.sample_theta_variational(np.tile(phi_tr[None], (S, 1, 1)), q_mu, q_sqrt, z=Z)
    np.testing.assert_allclose(Theta, theta_ref[:, :, 0].T, rtol=1e-9, atol=1e-12)
    X = rng.uniform(-1, 1, (N, d))
    F = rff_oracle.rff_values_shared(X, W, b, Theta, chunk=64)
    fvals, idx, _ = rff_oracle.eval_argmax_gather(np.tile(X[None], (S, 1, 1)), Wc, bc, theta_ref)
    np.testing.assert_allclose(F, fvals.T, rtol=1e-10, atol=1e-12)
    val, aidx, second = rff_oracle.rff_argmax_shared(X, W, b, Theta, chunk=64, idx_offset=1000)
    np.testing.assert_array_equal(aidx - 1000, idx)
    np.testing.assert_allclose(val, fvals.max(axis=1), rtol=1e-12)
    np.testing.assert_allclose(second, np.sort(fvals, axis=1)[:, -2], rtol=1e-12)


def test_rmsprop_ascent_improves_and_clips():
    rng = np.random.default_rng(3)
    count, n_init, d, D = 3, 5, 1, 32
    W = rng.standard_normal((count, D, d)) / 0.2
    b = rng.uniform(0, 2 * np.pi, (count, D, 1))
    theta = rng.standard_normal((count, D, 1))
    x0 = rng.uniform(0, 1, (count, n_init, d))
    x, steps = rff_oracle.rmsprop_rho0_ascent(x0, W, b, theta, 0.0, 1.0, num_steps=200)
    assert steps <= 200 and x.min() >= 0.0 and x.max() <= 1.0
    assert rff_oracle.maximizer_objective(x, W, b, theta) < rff_oracle.maximizer_objective(x0, W, b, theta)
    # analytic gradient vs central differences
    g = rff_oracle.maximizer_objective_grad(x0, W, b, theta)
    h = 1e-6
    xp, xm = x0.copy(), x0.copy()
    xp[1, 2, 0] += h
    xm[1, 2, 0] -= h
    fd = (rff_oracle.maximizer_objective(xp, W, b, theta) - rff_oracle.maximizer_objective(xm, W, b, theta)) / (2 * h)
    assert abs(fd - g[1, 2, 0]) < 1e-6 * max(1.0, abs(fd))
    # every step moves each coordinate by ~lr (sign-gradient behaviour of rho=0 RMSprop)
    x1, _ = rff_oracle.rmsprop_rho0_ascent(x0, W, b, theta, -10.0, 10.0, num_steps=1)
    assert np.all(np.abs(np.abs(x1 - x0) - 1e-3) < 2e-4)


@pytest.mark.parametrize("tag", ["forr", "shc", "stop"])
def test_ascent_loop_matches_reference_golden(golden, tag):
    """The restated Keras-RMSprop(rho=0) loop against the reference's own loop ff.py:139-157, executed unmodified under
    the shim's torch-autograd ``minimize`` (tests/golden/make_golden.py:golden_ascent): iterates, step count (early stop
    included) and the final argmax / gather."""
    g = golden("ff_ascent.npz")
    count, n_init, D, d, steps, steps_run = [int(v) for v in g[f"{tag}_meta"]]
    lo, hi = g[f"{tag}_bounds"]
    W, b, theta, x0 = g[f"{tag}_W"], g[f"{tag}_b"], g[f"{tag}_theta"], g[f"{tag}_x0"]
    assert x0.shape == (count, n_init, d)
    x, ran = rff_oracle.rmsprop_rho0_ascent(x0, W, b, theta, lo, hi, num_steps=steps)
    assert ran == steps_run
    if tag == "stop":
        assert steps_run < steps                      # the early-stop branch (ff.py:154-156) was taken
    # (g / (abs(g) + eps) amplifies the last-bit difference between the analytic and the autograd gradient where g ~ eps)
    np.testing.assert_allclose(x, g[f"{tag}_x_final"], rtol=0, atol=1e-11)
    _, _, maxi = rff_oracle.eval_argmax_gather(x, W, b, theta)
    np.testing.assert_allclose(maxi, g[f"{tag}_maximizers"], rtol=0, atol=1e-11)
    # analytic gradient of the restatement == autograd gradient the reference loop used: one step reproduces
    x1, _ = rff_oracle.rmsprop_rho0_ascent(x0, W, b, theta, lo, hi, num_steps=1)
    assert np.abs(x1 - x0).max() <= 1e-3 + 1e-12


def test_pes_grid_maximizer_samplers(golden):
    """pes.sample_maximizers_discrete / _simple (pes.py:9-41) restated: batches of exact draws, zero tail, argmax, take."""
    g = golden("pes_maximizers.npz")
    count, num_data, d, batch, n_disc = [int(v) for v in g["meta"]]
    data, batches = g["data"], g["batches"].astype(np.float64)
    np.testing.assert_array_equal(mpes_oracle.grid_maximizers_from_batches(data, batches, batch), g["maximizers"])
    np.testing.assert_array_equal(mpes_oracle.grid_maximizers_from_batches(data, batches - 10.0, batch), g["maximizers_neg"])
    xx = np.linspace(0.0, 1.0, n_disc).reshape(n_disc, 1)
    simple = g["simple_samples"].astype(np.float64)
    np.testing.assert_array_equal(xx[np.argmax(simple, axis=1)], g["maximizers_simple"])


def test_singular_retry_raises_failed():
    """ff.py:92-105: an always-singular regression exhausts 100 retries and raises ValueError("Failed")."""
    rng = np.random.default_rng(0)
    X = np.zeros((2, 4, 1))               # identical train points -> phi phi^T is rank 1
    with pytest.raises(ValueError, match="Failed"):
        rff_oracle.sample_features_weights(X, np.array([0.2]), np.zeros((4, 1)), np.eye(4)[None], 16, rng)


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not mounted (GPU box)")
def test_live_reference_agrees():
    """Re-run the reference sources now (fresh random inputs) and compare with the oracle."""
    ref = ref_loader.load()
    rng = np.random.default_rng(99)
    S, Q, C, k, M = 128, 6, 4, 2, 5
    fs = rng.standard_normal((S, Q * C + M))
    fs[0, Q * C + M - 1] = 10.0
    model = tf_shim.StubModel(fixed_samples=fs)
    with contextlib.redirect_stdout(io.StringIO()):
        mi_ref = ref.rank_pes.I_batch(rng.uniform(size=(Q, C, 2)), rng.uniform(size=(M, 2)), model, topk=k, num_samples=S)
    mi = mpes_oracle.rank_mpes_from_samples(fs[:, :Q * C].reshape(S, Q, C), fs[:, Q * C:], topk=k)
    np.testing.assert_allclose(mi, mi_ref, rtol=1e-9, atol=1e-13)
    X, W, b = rng.uniform(size=(2, 5, 3)), rng.standard_normal((2, 8, 3)), rng.uniform(size=(2, 8, 1))
    np.testing.assert_allclose(rff_oracle.fourier_features(X, W, b),
                               ref.fourier_features.fourier_features(tf_shim._t(X), tf_shim._t(W), tf_shim._t(b)).numpy(),
                               rtol=0, atol=1e-15)
