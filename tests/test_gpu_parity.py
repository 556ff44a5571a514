"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every test calls the product through its public
Python surface, which reaches the kernels through the C ABI (ctypes -> libtkbo.so), and compares with the
fp64 CPU oracle and/or the reference-produced golden vectors in tests/golden.

Tolerances (BASELINE.json north star): sampled function values within 1e-4 relative to the sample's scale
(max |f| over the candidates); maximiser indices identical except where the fp64 top-2 gap is below that
tolerance (tie counts are asserted small and printed); acquisition values within 1e-4 relative.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import mpes_oracle, rff_oracle

pytestmark = pytest.mark.gpu

VAL_RTOL = 1e-4


def _basis_inputs(rng, N, D, d, S, lo, hi, ls, realistic_theta=True, n_train=24):
    X = rng.uniform(lo, hi, size=(N, d)).astype(np.float32)
    W = rng.standard_normal((D, d)) / ls
    b = rng.uniform(0, 2 * np.pi, D)
    if realistic_theta:
        # theta from the ff.py:56-87 regression on a synthetic fitted posterior
        Xtr = rng.uniform(lo, hi, size=(n_train, d))
        K = np.exp(-0.5 * (((Xtr[:, None] - Xtr[None]) / ls) ** 2).sum(-1))
        L = np.linalg.cholesky(K + 1e-6 * np.eye(n_train))
        q_mu, q_sqrt = L @ rng.standard_normal((n_train, 1)), (0.3 * L)[None]
        phi = rff_oracle.fourier_features(Xtr[None], W[None], b[None, :, None])[0]
        Theta = rff_oracle.theta_shared_basis(phi, q_mu, q_sqrt, rng.standard_normal((S, n_train)))
    else:
        Theta = rng.standard_normal((D, S)) * rng.uniform(0.05, 20.0, size=(1, S))
    return X, W, b, Theta


def _check_argmax(val, idx, X, W, b, Theta, idx_offset=0):
    ref_val, ref_idx, second, absmax = rff_oracle.rff_argmax_shared(X.astype(np.float64), W, b, Theta, return_absmax=True)
    F_scale = np.maximum(absmax, 1e-300)
    val = val.cpu().numpy().astype(np.float64)
    idx = idx.cpu().numpy() - idx_offset
    assert np.all((idx >= 0) & (idx < X.shape[0]))
    assert np.max(np.abs(val - ref_val) / F_scale) < VAL_RTOL
    mism = idx != ref_idx
    gap = (ref_val - second) / F_scale
    ties = int(mism.sum())
    assert not np.any(mism & (gap > VAL_RTOL)), "argmax differs where the top-2 gap exceeds tolerance"
    return ties


@pytest.mark.parametrize("N,D,d,S,lo,hi,ls", [
    (128, 64, 2, 32, -1.5, 1.5, 0.6),         # one tile, one slab
    (1000, 200, 1, 20, 0.0, 1.0, 0.2),        # ragged N, D not a multiple of 64, S not a multiple of 32
    (5000, 1000, 6, 96, 0.0, 1.0, 0.2),       # SUSHI-like d=6
    (4096, 256, 4, 256, 0.0, 1.0, 0.3),       # single-buffered 256-sample accumulator
    (3000, 192, 3, 300, -1.0, 1.0, 0.5),      # two sample chunks of 160
    (2048, 128, 12, 512, 0.0, 1.0, 0.8),      # d=12 (SUSHI duels), two chunks of 256
    (777, 70, 5, 7, 0.0, 1.0, 0.4),           # d padded to the 6-wide template, tiny S
    (20000, 512, 2, 64, -1.5, 1.5, 0.6),      # many tiles per CTA
    (40000, 64, 2, 64, -1.5, 1.5, 0.6),       # one slab per item: candidate rows staged every slab, 2 items per CTA
    (40000, 256, 4, 256, 0.0, 1.0, 0.3),      # odd smem ring (3 stages): single projection warp, several items per CTA
    (6000, 320, 12, 256, 0.0, 1.0, 0.8),      # K = 75 projection (two W tiles), single candidate-row buffer
    (30000, 128, 16, 32, 0.0, 1.0, 1.0),      # d = 16 (K = 99), 16-column accumulator halves
    (9000, 192, 8, 40, 0.0, 1.0, 0.7),        # d = 8 (K = 51, one 128-byte W tile)
    (9000, 192, 3, 130, -1.0, 1.0, 0.5),      # d = 3 (K = 21, 64-byte swizzle), S chunk of 160
    (31, 1, 12, 128, 0.0, 1.0, 0.8),          # one slab, one row buffer, two projection warps: the second owns nothing
    (40000, 40, 12, 128, 0.0, 1.0, 0.8),      # same, 3 items per CTA: each projection warp skips every other item
])
def test_rff_eval_and_argmax_vs_oracle(pbo, N, D, d, S, lo, hi, ls):
    rng = np.random.default_rng(N + D + S)
    # 1-D inputs with 24 training points make phi phi^T numerically singular (cond ~ 1e16): see
    # test_rff_error_scales_with_weight_norm for that regime; here keep the regression well posed.
    X, W, b, Theta = _basis_inputs(rng, N, D, d, S, lo, hi, ls, n_train=8 if d == 1 else 24)
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    F = basis.eval(X).cpu().numpy().T.astype(np.float64)
    Fref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)
    scale = np.abs(Fref).max(axis=0)
    assert np.isfinite(F).all()
    assert np.max(np.abs(F - Fref) / scale) < VAL_RTOL
    val, idx = basis.argmax(X, idx_offset=12345)
    _check_argmax(val, idx, X, W, b, Theta, idx_offset=12345)
    # calling again gives identical results (workspace is re-initialised per call; the two MMA-issuing warps keep
    # the accumulation order fixed), values included
    for _ in range(3):
        val2, idx2 = basis.argmax(X, idx_offset=12345)
        assert torch.equal(val, val2) and torch.equal(idx, idx2)
    assert torch.equal(basis.eval(X), basis.eval(X))


@pytest.mark.parametrize("N,D,d,S", [(20000, 512, 6, 256), (9000, 320, 4, 300), (30000, 256, 2, 64), (5000, 192, 12, 128)])
def test_rff_precision_modes(pbo, N, D, d, S):
    """"fp32" (three fp16 MMAs per product) holds ~1e-5 of max|f|; "fast" (fp16 hi*hi + one e4m3 MMA carrying both
    correction terms) holds the path's 1e-4; "auto" without a scale for f is fp32 (it never trades accuracy blindly); with
    ``f_scale`` it takes fast only when the bound 1e-4 ||theta'|| fits in 0.3e-4 max|f|.  Theta columns span a
    wide dynamic range inside each sample (the e4m3 tile keeps 8 octaves below the column maximum in its normal range)."""
    rng = np.random.default_rng(N + S)
    X, W, b, Theta = _basis_inputs(rng, N, D, d, S, 0.0, 1.0, 0.4)
    Theta = Theta * np.exp(2.0 * rng.standard_normal((D, 1)))
    Fref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)
    scale = np.abs(Fref).max(axis=0)
    errs = {}
    for precision in ("fp32", "fast", "auto"):
        basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision=precision)
        cfg = basis.config(N)
        assert cfg["e4m3_corrections"] == {"fp32": 0, "fast": 1, "auto": 0}[precision]
        assert basis.precision == {"fp32": "fp32", "fast": "fast", "auto": "fp32"}[precision]
        F = basis.eval(X).cpu().numpy().T.astype(np.float64)
        errs[precision] = float(np.max(np.abs(F - Fref) / scale))
        val, idx = basis.argmax(X)
        _check_argmax(val, idx, X, W, b, Theta)
        val2, idx2 = basis.argmax(X)
        assert torch.equal(val, val2) and torch.equal(idx, idx2)
    assert errs["fp32"] < 1e-5 and errs["fast"] < VAL_RTOL, errs
    assert errs["fast"] > errs["fp32"]                     # the modes really differ
    # auto with a scale: fast only if provably inside the tolerance and only in the tensor-bound regime (S > 96)
    FF = pbo.fourier_features
    norm = np.linalg.norm(Theta * np.sqrt(2.0 / D), axis=0)
    generous = FF.SharedBasis(W, b, Theta, device="cuda:0", precision="auto", f_scale=norm * 10.0)      # bound = 1e-5 max|f|
    tight = FF.SharedBasis(W, b, Theta, device="cuda:0", precision="auto", f_scale=norm)                # bound = 1e-4 max|f|
    assert generous.precision == ("fast" if S > 96 else "fp32") and tight.precision == "fp32"
    assert float(generous.error_bound().max()) <= 0.3e-4 * float(norm.max() * 10.0) or generous.precision == "fp32"
    # set() re-decides the form for new weights and rebuilds the layout when it changes
    tight.set(W, b, Theta, f_scale=norm * 10.0)
    assert tight.precision == generous.precision
    np.testing.assert_array_equal(tight.eval(X).cpu().numpy(), generous.eval(X).cpu().numpy())


@pytest.mark.parametrize("N,D,d,S", [(70000, 512, 6, 256), (33000, 384, 4, 700), (129, 128, 2, 100), (50000, 256, 12, 160)])
def test_cta_pairs_equal_single_ctas(pbo, N, D, d, S, monkeypatch):
    """CTA pairs (cta_group::2: one M = 256 MMA per pair, operands fetched half by each CTA, second epilogue warp set) are
    a different schedule of the same arithmetic: per accumulator column the same MMAs in the same order, so argmax, values
    and stored samples are BITWISE those of the single-CTA launch - for odd tile counts (the peer's last tile is empty),
    several sample chunks and both contraction forms."""
    rng = np.random.default_rng(N)
    X, W, b, Theta = _basis_inputs(rng, N, D, d, S, 0.0, 1.0, 0.3)
    Xd = torch.from_numpy(X).cuda()
    for form in ("fp32", "fast"):
        res = {}
        for pair in ("0", "1"):
            monkeypatch.setenv("TKBO_PAIR", pair)
            basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision=form)
            assert basis.config(N)["cta_pairs"] == int(pair)
            res[pair] = (basis.argmax(Xd), basis.eval(Xd[:3000]))
        (v0, i0), F0 = res["0"]
        (v1, i1), F1 = res["1"]
        assert torch.equal(v0, v1) and torch.equal(i0, i1) and torch.equal(F0, F1)
    _check_argmax(v1, i1, X, W, b, Theta)


def test_peer_merge_single_rank(pbo):
    """The peer-memory merge with one rank (its own buffer is the only peer): keys pushed by the kernel's last CTA, flags,
    epoch parity and the merge kernel give the same (value, index) as the plain call, over several epochs and shapes of
    the same buffer.  (Two and eight ranks: tools/peer_merge_check.py under torchrun.)"""
    rng = np.random.default_rng(77)
    X, W, b, Theta = _basis_inputs(rng, 5000, 192, 3, 40, 0.0, 1.0, 0.5)
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    pm = pbo.distributed.PeerMerge(basis.S, "cuda:0")
    for it in range(5):
        Xi = np.ascontiguousarray(X[it * 700:(it + 1) * 700 + 1500])
        val, idx = pm.argmax(basis, Xi, idx_offset=1000 * it)
        val0, idx0 = basis.argmax(Xi, idx_offset=1000 * it)
        assert torch.equal(val, val0) and torch.equal(idx, idx0)
    assert pm.epoch == 5
    # an empty shard delivers "no candidate": value NaN, index -1, and the next epoch works as before
    val, idx = pm.argmax(basis, torch.empty((0, 3), device="cuda:0"), idx_offset=0)
    assert bool(torch.isnan(val).all()) and bool((idx == -1).all())
    val, idx = pm.argmax(basis, X[:900], idx_offset=7)
    val0, idx0 = basis.argmax(X[:900], idx_offset=7)
    assert torch.equal(val, val0) and torch.equal(idx, idx0)
    # delivery + separate merge launch (fused=False) gives the same bits as the merge in the kernel's last CTA
    val, idx = pm.argmax(basis, X[:900], idx_offset=7, fused=False)
    assert torch.equal(val, val0) and torch.equal(idx, idx0)
    torch.cuda.synchronize()
    pm.check()                                             # no merge gave up
    # the host-buffer form of the same (tkbo_rff_argmax_host_peer): one launch streams the shard in, delivers, merges
    hv, hi = pbo.fourier_features.argmax_host(W, b, Theta, X[:4000], device=0, peer=pm, idx_offset=11)
    val0, idx0 = basis.argmax(X[:4000], idx_offset=11)
    np.testing.assert_array_equal(hv, val0.cpu().numpy())
    np.testing.assert_array_equal(hi, idx0.cpu().numpy())
    # a peer that never delivers: the merge gives up after its bounded wait, reports index -2 and the object goes dead
    nat = pbo._native
    import ctypes
    lonely = pbo.distributed.PeerMerge(basis.S, "cuda:0")
    lonely.world, lonely.rank = 2, 0                       # pretend there is a second rank that stays silent
    words = int(nat.lib().tkbo_peer_buffer_bytes(2, basis.S)) // 8
    lonely.buf = torch.zeros(words, dtype=torch.int64, device="cuda:0")
    lonely._ptrs = (ctypes.c_void_p * 2)(lonely.buf.data_ptr(), lonely.buf.data_ptr())
    val, idx = lonely.argmax(basis, X[:900], idx_offset=0)
    torch.cuda.synchronize()
    assert bool((idx == -2).all())
    with pytest.raises(nat.TkboError, match="did not deliver"):
        lonely.check()
    with pytest.raises(nat.TkboError, match="timed out"):
        lonely.argmax(basis, X[:900], idx_offset=0)


def test_rff_argmax_wide_dynamic_range_theta(pbo):
    """Per-sample power-of-two scaling keeps tiny and huge weight samples equally accurate."""
    rng = np.random.default_rng(5)
    X, W, b, Theta = _basis_inputs(rng, 3000, 256, 2, 64, -1.5, 1.5, 0.6, realistic_theta=False)
    Theta[:, 0] *= 1e-12
    Theta[:, 1] *= 1e12
    Theta[:, 2] = 0.0                       # all-equal sample: exact ties everywhere -> index 0
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    val, idx = basis.argmax(X)
    _check_argmax(val, idx, X, W, b, Theta)
    assert int(idx[2]) == 0 and float(val[2]) == 0.0


def test_rff_error_scales_with_weight_norm(pbo):
    """The fp32-grade feature map bounds the absolute error by ~4e-6 * ||sqrt(2/D) theta_s||_2.  With an
    ill-posed regression (24 training points in 1-D, cond(phi phi^T) ~ 1e16, no jitter in ff.py:80-82) the
    weights cancel massively (||theta'|| / max|f| ~ 1e3) and the relative error grows by that factor; the
    bound reported by SharedBasis.error_bound() must still hold, and the argmax must stay within it."""
    rng = np.random.default_rng(1000 + 200 + 20)
    X, W, b, Theta = _basis_inputs(rng, 1000, 200, 1, 20, 0.0, 1.0, 0.2, n_train=24)
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    F = basis.eval(X).cpu().numpy().T.astype(np.float64)
    Fref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)
    bound = basis.error_bound().cpu().numpy()
    kappa = np.linalg.norm(Theta * np.sqrt(2.0 / 200), axis=0) / np.abs(Fref).max(axis=0)
    assert kappa.max() > 100                                     # this really is the ill-conditioned regime
    assert np.all(np.abs(F - Fref).max(axis=0) <= bound)
    val, idx = basis.argmax(X)
    ref_val = Fref.max(axis=0)
    got = Fref[idx.cpu().numpy(), np.arange(20)]                 # fp64 value at the returned index
    assert np.all(ref_val - got <= 2 * bound) and np.all(np.abs(val.cpu().numpy() - ref_val) <= bound)


def test_rff_argmax_exact_ties_pick_lowest_index(pbo):
    """Duplicate candidates produce bit-identical values; np.argmax / tf.argmax keep the first."""
    rng = np.random.default_rng(6)
    X, W, b, Theta = _basis_inputs(rng, 1024, 128, 2, 32, -1.5, 1.5, 0.6)
    ref_val, ref_idx, _ = rff_oracle.rff_argmax_shared(X.astype(np.float64), W, b, Theta)
    Xd = np.concatenate([X, X, X], axis=0)                     # every maximiser now appears 3 times, 1024 rows apart
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    _, idx = basis.argmax(Xd)
    idx = idx.cpu().numpy()
    assert np.all(idx < 1024)
    val1, idx1 = basis.argmax(X)
    np.testing.assert_array_equal(idx, idx1.cpu().numpy())


def test_rff_sharded_equals_single(pbo):
    """Candidate shards with idx_offset + merge == one call over all candidates (the multi-GPU decomposition)."""
    rng = np.random.default_rng(7)
    X, W, b, Theta = _basis_inputs(rng, 10000, 256, 3, 48, 0.0, 1.0, 0.3)
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    val, idx = basis.argmax(X)
    vals, idxs = [], []
    for r in range(4):
        lo, hi = pbo.distributed.shard_range(10000, r, 4)
        v, i = basis.argmax(X[lo:hi], idx_offset=lo)
        vals.append(v)
        idxs.append(i)
    mv, mi = pbo.distributed.merge_pairs(torch.stack(vals).cpu(), torch.stack(idxs).cpu())
    assert torch.equal(mv, val.cpu()) and torch.equal(mi, idx.cpu())
    # and the C-ABI merge kernel agrees
    nat = pbo._native
    ov = torch.empty(48, dtype=torch.float32, device="cuda:0")
    oi = torch.empty(48, dtype=torch.int64, device="cuda:0")
    sv, si = torch.stack(vals).contiguous(), torch.stack(idxs).contiguous()
    nat.check(nat.lib().tkbo_merge_argmax(nat.handle(0), nat.ptr(sv), nat.ptr(si), 4, 48, nat.ptr(ov), nat.ptr(oi),
                                          nat.current_stream_ptr(torch.device("cuda:0"))), "merge")
    assert torch.equal(ov, val) and torch.equal(oi, idx)


def test_rff_argmax_key_form_and_merge(pbo):
    """The int64 merge-key output: max over shard keys == the single-call (value, lowest index), including an exact
    cross-shard tie (duplicated candidates) and an empty shard."""
    rng = np.random.default_rng(17)
    X, W, b, Theta = _basis_inputs(rng, 6000, 128, 2, 40, -1.5, 1.5, 0.6)
    X = np.concatenate([X, X[:3000]], axis=0)              # rows 6000.. duplicate rows 0..2999 -> exact ties
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    val, idx = basis.argmax(X)
    keys = []
    for r in range(3):
        lo, hi = pbo.distributed.shard_range(X.shape[0], r, 3)
        keys.append(basis.argmax_key(X[lo:hi], idx_offset=lo))
    keys.append(torch.full((40,), pbo.distributed.EMPTY_KEY, dtype=torch.int64, device="cuda:0"))
    merged = torch.stack(keys).max(dim=0).values
    v2, i2 = basis.unpack_key(merged)
    assert torch.equal(v2, val) and torch.equal(i2, idx)
    # the torch restatement of the key (used by the gloo tests of the merge) is bit-identical to the device's
    assert torch.equal(pbo.distributed.pack_argmax_key(val, idx), merged)
    v2b, i2b = pbo.distributed.unpack_argmax_key(merged)
    assert torch.equal(v2b, val) and torch.equal(i2b, idx)
    assert int(idx.max()) < 6000                           # ties resolved to the first copy
    ve, ie = basis.unpack_key(keys[-1])
    assert bool((ie == -1).all()) and bool(torch.isnan(ve).all())
    # single-process call of the distributed helper (no process group): identity merge
    v3, i3 = pbo.distributed.sharded_argmax_fast(basis, torch.from_numpy(X).cuda(), 0)
    assert torch.equal(v3, val) and torch.equal(i3, idx)


def test_rff_argmax_ignores_nan_candidates(pbo):
    """Candidates with a NaN coordinate produce NaN samples; they are never selected (np.nanargmax semantics), a
    sample whose candidates are all NaN returns (NaN, -1)."""
    rng = np.random.default_rng(23)
    X, W, b, Theta = _basis_inputs(rng, 3000, 128, 2, 40, -1.5, 1.5, 0.6)
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    _, idx0 = basis.argmax(X)
    Xn = X.copy()
    bad = np.unique(np.concatenate([idx0.cpu().numpy()[:20], np.arange(64, 128), rng.integers(0, 3000, 200)]))
    Xn[bad, rng.integers(0, 2, bad.shape[0])] = np.nan        # includes current winners and one whole 32-row warp slice
    val, idx = basis.argmax(Xn)
    keep = np.setdiff1d(np.arange(3000), bad)
    ref_val, ref_idx, second, absmax = rff_oracle.rff_argmax_shared(X[keep].astype(np.float64), W, b, Theta, return_absmax=True)
    got = idx.cpu().numpy()
    assert not np.isin(got, bad).any() and np.isfinite(val.cpu().numpy()).all()
    mism = got != keep[ref_idx]
    assert not np.any(mism & ((ref_val - second) / absmax > VAL_RTOL))
    assert np.max(np.abs(val.cpu().numpy() - ref_val) / absmax) < VAL_RTOL
    v2, i2 = basis.argmax(np.full((300, 2), np.nan, dtype=np.float32))
    assert bool((i2 == -1).all()) and bool(torch.isnan(v2).all())


def test_rff_argmax_host_entry_point(pbo):
    """tkbo_rff_argmax_host: the pure C-ABI end-to-end call with host buffers (no torch memory involved)."""
    rng = np.random.default_rng(8)
    X, W, b, Theta = _basis_inputs(rng, 2500, 128, 2, 40, -1.5, 1.5, 0.6)
    nat = pbo._native
    val = np.empty(40, dtype=np.float32)
    idx = np.empty(40, dtype=np.int64)
    as_p = lambda a: a.ctypes.data_as(ctypes.c_void_p)   # noqa: E731
    Wc, bc, Tc = np.ascontiguousarray(W), np.ascontiguousarray(b), np.ascontiguousarray(Theta)
    nat.check(nat.lib().tkbo_rff_argmax_host(nat.handle(0), as_p(Wc), as_p(bc), as_p(Tc), 128, 2, 40, as_p(X), 2500, 0,
                                             as_p(val), as_p(idx)), "argmax_host")
    _check_argmax(torch.from_numpy(val), torch.from_numpy(idx), X, W, b, Theta)
    # the CUDA_LAUNCH_BLOCKING / profiler ordering (every copy queued before the launch) gives the same result
    import os
    os.environ["TKBO_HOST_STREAMED"] = "0"
    try:
        val2, idx2 = np.empty(40, dtype=np.float32), np.empty(40, dtype=np.int64)
        nat.check(nat.lib().tkbo_rff_argmax_host(nat.handle(0), as_p(Wc), as_p(bc), as_p(Tc), 128, 2, 40, as_p(X), 2500, 0,
                                                 as_p(val2), as_p(idx2)), "argmax_host")
    finally:
        del os.environ["TKBO_HOST_STREAMED"]
    np.testing.assert_array_equal(val, val2)
    np.testing.assert_array_equal(idx, idx2)


def test_rff_argmax_host_chunked_pipeline(pbo):
    """Large N goes through the streamed pipeline of the host entry point (one launch that consumes candidate chunks as
    their H2D copies land): same result as the device-resident call, with pinned and with pageable candidates (whose
    copies are queued, host-synchronously, after the kernel is already waiting for them); repeated calls reuse the
    cached staging, a shape change rebuilds it."""
    rng = np.random.default_rng(9)
    X, W, b, Theta = _basis_inputs(rng, 400000, 128, 2, 40, -1.5, 1.5, 0.6)
    Xp = torch.from_numpy(X).pin_memory()
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    val, idx = basis.argmax(X)
    for src in (Xp, Xp, X, Xp):
        v, i = pbo.fourier_features.argmax_host(W, b, Theta, src)
        np.testing.assert_array_equal(i, idx.cpu().numpy())
        np.testing.assert_array_equal(v, val.cpu().numpy())
    # the winners sit in late chunks too: a kernel that read a chunk before it landed would miss them
    assert (idx.cpu().numpy() > 200000).any()
    # explicit fast form through the host path == the device-resident fast form
    fast = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision="fast")
    vf, if_ = fast.argmax(X)
    v, i = pbo.fourier_features.argmax_host(W, b, Theta, Xp, precision="fast")
    np.testing.assert_array_equal(i, if_.cpu().numpy())
    np.testing.assert_array_equal(v, vf.cpu().numpy())
    X2, W2, b2, Theta2 = _basis_inputs(rng, 3000, 64, 3, 20, 0.0, 1.0, 0.4)
    v, i = pbo.fourier_features.argmax_host(W2, b2, Theta2, X2)
    val2, idx2 = pbo.fourier_features.SharedBasis(W2, b2, Theta2, device="cuda:0").argmax(X2)
    np.testing.assert_array_equal(i, idx2.cpu().numpy())
    np.testing.assert_array_equal(v, val2.cpu().numpy())


def test_argument_validation_errors(pbo):
    nat = pbo._native
    basis = ctypes.c_void_p()
    assert nat.lib().tkbo_rff_basis_create(nat.handle(0), ctypes.byref(basis), 64, 40, 8) == -3   # d > 16 unsupported
    assert "not supported" in nat.last_error()
    assert nat.lib().tkbo_rff_basis_create(nat.handle(0), ctypes.byref(basis), 0, 2, 8) == -1
    nat.check(nat.lib().tkbo_rff_basis_create(nat.handle(0), ctypes.byref(basis), 64, 2, 8), "create")
    x = torch.zeros((4, 2), device="cuda:0")
    v = torch.zeros(8, device="cuda:0")
    i = torch.zeros(8, dtype=torch.int64, device="cuda:0")
    rc = nat.lib().tkbo_rff_argmax(nat.handle(0), basis, nat.ptr(x), 4, 0, nat.ptr(v), nat.ptr(i), None)
    assert rc == -1 and "basis_set" in nat.last_error()                                        # unset basis
    nat.lib().tkbo_rff_basis_destroy(nat.handle(0), basis)


# ------------------------------------------------------------------------------------------- reference-shaped API
@pytest.mark.parametrize("tag", ["forr", "shc"])
def test_eval_argmax_gather_matches_reference_golden(pbo, golden, tag):
    """ff.py:162-167 on the exact arrays the reference drew (per-sample basis, float64)."""
    g = golden("ff_sample_maximizers.npz")
    fvals, idx, maxi = pbo.fourier_features.eval_argmax_gather(g[f"{tag}_x_star"], g[f"{tag}_W"], g[f"{tag}_b"],
                                                               g[f"{tag}_theta"], device="cuda:0")
    np.testing.assert_array_equal(maxi.cpu().numpy(), g[f"{tag}_maximizers"])
    f_ref, idx_ref, _ = rff_oracle.eval_argmax_gather(g[f"{tag}_x_star"], g[f"{tag}_W"], g[f"{tag}_b"], g[f"{tag}_theta"])
    np.testing.assert_allclose(fvals.cpu().numpy(), f_ref, rtol=0, atol=1e-11 * np.abs(f_ref).max())
    np.testing.assert_array_equal(idx.cpu().numpy(), idx_ref)


@pytest.mark.parametrize("tag", ["under", "over"])
def test_fourier_features_and_theta_match_reference_golden(pbo, golden, tag):
    g = golden("ff_features_theta.npz")
    ff = pbo.fourier_features
    phi = ff.fourier_features(g[f"{tag}_X"], g[f"{tag}_W"], g[f"{tag}_b"], device="cuda:0")
    np.testing.assert_allclose(phi.cpu().numpy(), g[f"{tag}_phi"], rtol=0, atol=1e-14)
    theta = ff.sample_theta_variational(phi, g[f"{tag}_q_mu"], g[f"{tag}_q_sqrt"], z=g[f"{tag}_z"])
    ref = g[f"{tag}_theta"]
    assert tuple(theta.shape) == ref.shape
    np.testing.assert_allclose(theta.cpu().numpy(), ref, rtol=0, atol=1e-6 * np.abs(ref).max())


def test_ascent_matches_oracle_restatement(pbo):
    """ff.py:139-157 on the device vs the NumPy restatement of Keras RMSprop(rho=0)."""
    rng = np.random.default_rng(9)
    count, n_init, D, d = 6, 10, 128, 2
    W = rng.standard_normal((count, D, d)) / 0.4
    b = rng.uniform(0, 2 * np.pi, (count, D, 1))
    theta = rng.standard_normal((count, D, 1))
    x0 = rng.uniform(-1, 1, (count, n_init, d))
    x_ref, steps_ref = rff_oracle.rmsprop_rho0_ascent(x0, W, b, theta, -1.0, 1.0, num_steps=400)
    x, steps = pbo.fourier_features.ascend(x0, W, b, theta, -1.0, 1.0, num_steps=400, device="cuda:0")
    assert steps == steps_ref
    # sign-gradient steps of size lr * g / (|g| + 1e-7): coordinates whose gradient passes through ~1e-7 amplify
    # the last-bit differences of the two summation orders, hence 1e-7 rather than 1e-12
    np.testing.assert_allclose(x.cpu().numpy(), x_ref, rtol=0, atol=1e-7)
    assert float(x.min()) >= -1.0 and float(x.max()) <= 1.0
    # early stop: a flat objective (theta = 0 would divide by zero like the reference; use tiny gradient instead)
    x1, steps1 = pbo.fourier_features.ascend(x0, W, b, theta, -1.0, 1.0, num_steps=0, device="cuda:0")
    assert steps1 == 0 and np.array_equal(x1.cpu().numpy(), x0)


@pytest.mark.parametrize("tag", ["forr", "shc", "stop"])
def test_ascent_matches_reference_golden(pbo, golden, tag):
    """tkbo_rff_ascent_batched against the iterates of the reference's own loop ff.py:139-157 (run unmodified under the
    shim's autograd RMSprop when the golden file was made): same step count - including the early stop - same final
    points to 1e-7, and the same maximisers after the final eval / argmax / gather (ff.py:162-167)."""
    g = golden("ff_ascent.npz")
    count, n_init, D, d, steps, steps_run = [int(v) for v in g[f"{tag}_meta"]]
    lo, hi = (float(v) for v in g[f"{tag}_bounds"])
    ff = pbo.fourier_features
    W, b, theta, x0 = g[f"{tag}_W"], g[f"{tag}_b"], g[f"{tag}_theta"], g[f"{tag}_x0"]
    x, ran = ff.ascend(x0, W, b, theta, lo, hi, num_steps=steps, device="cuda:0")
    assert ran == steps_run
    np.testing.assert_allclose(x.cpu().numpy(), g[f"{tag}_x_final"], rtol=0, atol=1e-7)
    _, idx, maxi = ff.eval_argmax_gather(x, W, b, theta, device="cuda:0")
    np.testing.assert_allclose(maxi.cpu().numpy(), g[f"{tag}_maximizers"], rtol=0, atol=1e-7)


def test_pes_grid_maximizer_samplers_match_reference_golden(pbo, golden):
    """pes.sample_maximizers_discrete / _simple with the reference's sampling scheme (exact draws per 1000-point batch
    through model.predict_f_samples, zero tail) against what the unmodified pes.py:9-41 returned for the same draws."""
    g = golden("pes_maximizers.npz")
    count, num_data, d, batch, n_disc = [int(v) for v in g["meta"]]

    class Stub:
        def __init__(self, batches):
            self.batches, self.calls = batches, 0

        def predict_f_samples(self, x, n):
            assert n == count and x.shape[0] == self.batches[self.calls].shape[1]
            out = torch.from_numpy(self.batches[self.calls].astype(np.float64)[:, :, None])
            self.calls += 1
            return out

    pes = pbo.acquisitions.pes
    batches = g["batches"]
    m = Stub(batches)
    maxi = pes.sample_maximizers_discrete(m, count, g["data"], batch_size=batch, device="cuda:0")
    assert m.calls == num_data // batch and tuple(maxi.shape) == (count, d) and maxi.dtype == torch.float64
    np.testing.assert_array_equal(maxi.numpy(), g["maximizers"])
    maxi_neg = pes.sample_maximizers_discrete(Stub(batches - 10.0), count, g["data"], batch_size=batch, device="cuda:0")
    np.testing.assert_array_equal(maxi_neg.numpy(), g["maximizers_neg"])           # the zero tail wins, as in the reference
    simple = pes.sample_maximizers_simple(Stub(g["simple_samples"][None]), count, n_disc, device="cuda:0")
    np.testing.assert_array_equal(simple.numpy(), g["maximizers_simple"])


def test_pes_grid_maximizer_sampler_rff_variant(pbo):
    """rff_features=D: RFF draws on one shared basis through the fused argmax kernel; the returned points are the oracle's
    arg max of the drawn functions (ties below tolerance excepted)."""
    mdl = pbo.model.synthetic_model(16, 2, 0.0, 1.0, 0.3, seed=5)
    rng = np.random.default_rng(6)
    data = rng.uniform(0, 1, size=(5000, 2))
    pbo.fourier_features.seed(11)
    maxi, info = pbo.acquisitions.pes.sample_maximizers_discrete(mdl, 24, data, rff_features=256, device="cuda:0",
                                                                 return_info=True)
    bs = info["basis"]
    ties = _check_argmax(info["values"], info["indices"], data.astype(np.float32), bs.W.cpu().numpy(), bs.b.cpu().numpy(),
                         bs.Theta.cpu().numpy())
    assert ties <= 2 and tuple(maxi.shape) == (24, 2)
    np.testing.assert_array_equal(maxi.numpy(), data[info["indices"].numpy()])


def test_sample_maximizers_reference_semantics(pbo):
    """End-to-end ff.py:110-169 at MPES_Forr defaults scaled down: shapes, bounds, and the returned point is the
    argmax start of each sample (checked with the oracle on the arrays the call actually drew)."""
    ff = pbo.fourier_features
    mdl = pbo.model.synthetic_model(10, 1, 0.0, 1.0, 0.2, seed=1)
    ff.seed(123)
    maxi, info = ff.sample_maximizers(X=mdl.Z, count=5, n_init=12, D=200, model=mdl, min_val=0.0, max_val=1.0,
                                      num_steps=60, device="cuda:0", return_info=True)
    assert tuple(maxi.shape) == (5, 1) and maxi.dtype == torch.float64 and maxi.device.type == "cpu"
    assert float(maxi.min()) >= 0.0 and float(maxi.max()) <= 1.0
    W, b, theta = (info[k].cpu().numpy() for k in ("W", "b", "theta"))
    _, _, ref_max = rff_oracle.eval_argmax_gather(info["x_star"].numpy(), W, b, theta)
    np.testing.assert_array_equal(maxi.numpy(), ref_max)
    # interpolation property of ff.py:80-84 (n < D): Phi(X_train) theta reproduces the q-samples it was fitted to
    assert theta.shape == (5, 200, 1) and info["steps"] <= 60


def test_sample_maximizers_grid_mode(pbo):
    ff = pbo.fourier_features
    mdl = pbo.model.synthetic_model(20, 2, -1.5, 1.5, 0.6, seed=2)
    grid = pbo.acquisitions.dts.uniform_grid(2, 64, -1.5, 1.5)
    ff.seed(7)
    maxi, info = ff.sample_maximizers(X=mdl.Z, count=16, n_init=50, D=256, model=mdl, min_val=-1.5, max_val=1.5,
                                      candidates=grid, device="cuda:0", return_info=True)
    assert tuple(maxi.shape) == (16, 2)
    bs = info["basis"]
    ties = _check_argmax(info["values"], info["indices"], grid.astype(np.float32), bs.W.cpu().numpy(), bs.b.cpu().numpy(),
                         bs.Theta.cpu().numpy())
    assert ties <= 2
    np.testing.assert_array_equal(maxi.numpy(), grid[info["indices"].numpy()])


def test_singular_regression_retries_then_fails(pbo):
    """ff.py:92-105: a regression that cannot be inverted on any draw exhausts 100 retries -> ValueError("Failed")."""
    ff = pbo.fourier_features
    # exactly singular Gram matrix [[1, 1], [1, 1]]: the LU hits a zero pivot (tf.linalg.inv -> InvalidArgumentError)
    phi = torch.tensor([[[1.0, 0.0, 0.0], [1.0, 0.0, 0.0]]], dtype=torch.float64, device="cuda:0")
    with pytest.raises(ff.SingularMatrixError):
        ff.sample_theta_variational(phi, np.zeros((2, 1)), np.eye(2)[None])
    # a zero lengthscale makes every draw of W infinite, so every retry fails the same way
    mdl = pbo.model.PreferentialGP(pbo.model.RBF([0.0]), np.zeros((4, 1)), np.zeros((4, 1)), np.eye(4)[None])
    X = np.zeros((2, 4, 1))
    import contextlib
    import io
    out = io.StringIO()
    with contextlib.redirect_stdout(out), np.errstate(all="ignore"), pytest.raises(ValueError, match="Failed"):
        ff.sample_features_weights(X, mdl, 16, device="cuda:0")
    assert out.getvalue().count("Resampling phi, W, b, theta") == 100 and "Retry limit exceeded" in out.getvalue()


@pytest.mark.parametrize("N,d,D,S,form", [(40000, 3, 192, 256, "fp32"), (40000, 3, 192, 256, "fast"), (30001, 6, 320, 300, "fast"),
                                          (9000, 4, 128, 160, "fp32")])
def test_two_pass_argmax_equals_single_pass(pbo, monkeypatch, N, d, D, S, form):
    """Large argmax calls run a hi*hi-only pass over everything and the full contraction only over the items that can hold a
    maximum (csrc/rff.cu: launch_argmax).  Forced on here (TKBO_PRUNE=2) against the single pass (TKBO_PRUNE=0): the same
    bits, with exact ties across tiles (lowest index wins), a NaN row and a ragged last tile in the input."""
    rng = np.random.default_rng(N + D + S)
    W = rng.standard_normal((D, d)) / 0.3
    b = rng.uniform(0, 2 * np.pi, D)
    Theta = rng.standard_normal((D, S)) * rng.uniform(0.5, 2.0, size=(1, S))
    X = rng.uniform(0.0, 1.0, size=(N, d)).astype(np.float32)
    ref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)          # (N, S)
    top = int(np.argmax(ref[:, 0]))
    X[N - 7] = X[top]                                   # an exact tie in another tile: the lower index has to win
    X[min(top + 1, N - 8)] = X[top] if top + 1 < N - 8 else X[min(top + 1, N - 8)]
    X[123] = np.nan
    Xd = torch.from_numpy(X).cuda()
    out = {}
    for prune in ("0", "2"):
        monkeypatch.setenv("TKBO_PRUNE", prune)
        basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision=form)
        val, idx = basis.argmax(Xd)
        tp = basis.two_pass_items()
        assert (tp is None) == (prune == "0")
        if tp is not None:
            assert 1 <= tp[0] <= tp[1]
        out[prune] = (val.clone(), idx.clone())
    assert torch.equal(out["0"][0], out["2"][0]) and torch.equal(out["0"][1], out["2"][1])
    assert int(out["2"][1][0]) == min(top, N - 7)


# ------------------------------------------------------------------------------------------- MPES / PES / DTS
@pytest.mark.parametrize("tag", ["c5k3", "c4k4", "c3k1", "c5k2_sparse", "gp_small", "gp_tiny", "gp_c4k2", "spread"])
def test_rank_pes_matches_reference_golden(pbo, golden, tag, capsys):
    """rank_pes.I_batch (imported unmodified from the reference when the golden file was made) vs K2 on the
    same fp32 samples.  gp_*: smooth GP-like samples with MI between 1e-4 and 7e-3 (the regime SURVEY 7 calls typical,
    where 1e-4 RELATIVE means ~1e-8 absolute); spread: choices > 100 apart (per-stage renormalisation, rank_pes.py:166)."""
    g = golden("rank_pes.npz")
    S, Q, C, k, M = [int(v) for v in g[f"{tag}_meta"]]
    mi = pbo.acquisitions.rank_pes.I_batch(np.zeros((Q, C, 6)), np.zeros((M, 6)), None, topk=k, num_samples=S,
                                           f_samples=g[f"{tag}_fsample_all"], device="cuda:0")
    assert "all possible observations" in capsys.readouterr().out                   # rank_pes.py:65-69
    ref = g[f"{tag}_mi"]
    np.testing.assert_allclose(mi.numpy(), ref, rtol=1e-4, atol=1e-9)
    if tag in ("gp_small", "gp_tiny"):
        # the same through the explicit-table kernel (the reference's > 1000-permutation path)
        from importlib import import_module
        common = import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
        fs = torch.from_numpy(g[f"{tag}_fsample_all"]).cuda()
        fchi, fstar = fs[:, :Q * C].contiguous().reshape(S, Q, C), fs[:, Q * C:].contiguous()
        table = mpes_oracle.all_permutations_k_in_n(C, k)
        np.testing.assert_allclose(common.mpes_from_device_samples(fchi, fstar, k, 0, table).cpu().numpy(), ref,
                                   rtol=1e-4, atol=1e-9)
        # bitwise repeatable (the block-wide group sums are integer reductions), also for a ragged / unaligned batch
        a1 = common.mpes_from_device_samples(fchi, fstar, k, 0, None)
        a2 = common.mpes_from_device_samples(fchi, fstar, k, 0, None)
        assert torch.equal(a1, a2)
        odd = fchi[:, :Q - 3, :].contiguous()
        np.testing.assert_allclose(common.mpes_from_device_samples(odd, fstar, k, 0, None).cpu().numpy(), ref[:Q - 3],
                                   rtol=1e-4, atol=1e-9)


def test_mpes_skewed_groups_large_batch(pbo):
    """One maximiser wins most samples (the usual state late in a BO run) and several win none: every warp of a block takes
    an equal share of the samples whatever the group sizes; result against the oracle on a slice, bitwise repeatable."""
    rng = np.random.default_rng(8)
    S, Q, C, k, M = 1024, 4000, 5, 3, 20
    basis = rng.standard_normal((Q * C, 5)).astype(np.float32)
    fchi_h = (rng.standard_normal((S, 5)).astype(np.float32) @ basis.T * 0.4).reshape(S, Q, C)
    fstar_h = rng.standard_normal((S, M)).astype(np.float32)
    fstar_h[:, 0] += 2.0                                     # maximiser 0 wins ~3/4 of the samples
    fstar_h[:, 15:] -= 10.0                                  # the last five never win
    counts = np.bincount(np.argmax(fstar_h, axis=1), minlength=M)
    assert counts[0] > 0.6 * S and (counts[15:] == 0).all()
    common = pbo.acquisitions._common
    fchi, fstar = torch.from_numpy(fchi_h).cuda(), torch.from_numpy(fstar_h).cuda()
    mi = common.mpes_from_device_samples(fchi, fstar, k, pbo._native.MPES_RANK)
    assert torch.equal(mi, common.mpes_from_device_samples(fchi, fstar, k, pbo._native.MPES_RANK))
    sl = slice(1000, 1200)
    ref = mpes_oracle.rank_mpes_from_samples(fchi_h[:, sl].astype(np.float64), fstar_h.astype(np.float64), topk=k)
    np.testing.assert_allclose(mi[sl].cpu().numpy(), ref, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("C,k", [(2, 1), (2, 2), (3, 1), (3, 2), (3, 3), (4, 1), (4, 2), (4, 3), (4, 4), (5, 1), (5, 2), (5, 3),
                                 (6, 1), (6, 2)])
@pytest.mark.parametrize("Q", [77, 84])
def test_mpes_every_register_resident_shape_matches_oracle(pbo, C, k, Q):
    """Every (C, k) that has an all-orderings instantiation (k >= 2 takes the two-wide mirror-image walk, k = 1 the scalar
    one) against the oracle port of rank_pes.I_batch on smooth samples (small MI).  Q = 77: rows that are not multiples of
    16 bytes for most C (direct loads); Q = 84: 16-byte rows with a ragged last tile of 20 queries (staged copies that stop
    at the end of the row)."""
    rng = np.random.default_rng(100 * C + k)
    S, M = 256, 7
    basis = rng.standard_normal((Q * C + M, 4)).astype(np.float32)
    f_all = (rng.standard_normal((S, 4)).astype(np.float32) @ basis.T * 0.6).astype(np.float32)
    fchi_h, fstar_h = f_all[:, :Q * C].reshape(S, Q, C).copy(), f_all[:, Q * C:].copy()
    common = pbo.acquisitions._common
    mi = common.mpes_from_device_samples(torch.from_numpy(fchi_h).cuda(), torch.from_numpy(fstar_h).cuda(), k,
                                         pbo._native.MPES_RANK)
    ref = mpes_oracle.rank_mpes_from_samples(fchi_h.astype(np.float64), fstar_h.astype(np.float64), topk=k)
    np.testing.assert_allclose(mi.cpu().numpy(), ref, rtol=1e-4, atol=1e-9)


@pytest.mark.parametrize("C,k", [(2, 1), (3, 2)])
def test_mpes_short_rows_do_not_touch_their_neighbours(pbo, C, k):
    """Rows of fewer than 32 16-byte chunks (C < 4): the staged copy must neither read past the tile's row segment nor write
    into the next ring slot.  Everything outside the queries' own columns is NaN here, so any stray byte poisons the result;
    the slice is compared with the same queries evaluated on their own."""
    rng = np.random.default_rng(31 + C)
    S, Q, M, pad = 300, 96, 9, 64
    basis = rng.standard_normal((Q * C + M, 3)).astype(np.float32)
    f_all = (rng.standard_normal((S, 3)).astype(np.float32) @ basis.T * 0.7).astype(np.float32)
    fchi_h, fstar_h = f_all[:, :Q * C].reshape(S, Q, C).copy(), f_all[:, Q * C:].copy()
    common = pbo.acquisitions._common
    wide = torch.full((S, Q + pad, C), float("nan"), device="cuda")
    wide[:, :Q] = torch.from_numpy(fchi_h).cuda()
    fstar = torch.from_numpy(fstar_h).cuda()
    ref = mpes_oracle.rank_mpes_from_samples(fchi_h.astype(np.float64), fstar_h.astype(np.float64), topk=k)
    alone = common.mpes_from_device_samples(torch.from_numpy(fchi_h).cuda(), fstar, k, pbo._native.MPES_RANK).cpu().numpy()
    np.testing.assert_allclose(alone, ref, rtol=1e-4, atol=1e-9)
    padded = common.mpes_from_device_samples(wide, fstar, k, pbo._native.MPES_RANK).cpu().numpy()
    assert np.array_equal(padded[:Q], alone)               # (the NaN queries themselves are not specified)


def test_rank_pes_log_likelihood_matches_reference_golden(pbo, golden):
    """rank_pes.get_log_likelihood / get_log_likelihood_given_order (rank_pes.py:111-177) against the (P, Q) table the
    unmodified reference functions produced."""
    g = golden("rank_pes.npz")
    S, Q, C, k, M = [int(v) for v in g["c5k3_meta"]]
    fchi = g["c5k3_fsample_all"].astype(np.float64)[:, :Q * C].reshape(S, Q, C)
    rp = pbo.acquisitions.rank_pes
    perms = rp.get_all_permutation_k_in_n(C, k)
    np.testing.assert_array_equal(perms, g["perms_5_3"])
    ll = rp.get_log_likelihood(fchi, perms, normalizer=S, device="cuda:0")
    assert tuple(ll.shape) == (60, Q) and ll.dtype == torch.float64
    np.testing.assert_allclose(ll.numpy(), g["c5k3_loglik"], rtol=1e-10, atol=1e-11)
    one = rp.get_log_likelihood_given_order(fchi, perms[7], normalizer=S, device="cuda:0")
    np.testing.assert_allclose(one.cpu().numpy(), g["c5k3_loglik_order7"], rtol=1e-10, atol=1e-11)


def test_rank_pes_through_model_predict_f_samples(pbo):
    """The reference call path: I_batch asks the model for (S, N, 1) samples (rank_pes.py:28)."""
    rng = np.random.default_rng(3)
    S, Q, C, M, d = 256, 9, 4, 5, 2
    fs = rng.standard_normal((S, Q * C + M)).astype(np.float32)

    class Stub:
        def predict_f_samples(self, x, n):
            assert x.shape == (Q * C + M, d) and n == S
            return torch.from_numpy(fs.astype(np.float64)[:, :, None])

    chi, xs = rng.uniform(size=(Q, C, d)), rng.uniform(size=(M, d))
    mi = pbo.acquisitions.rank_pes.I_batch(chi, xs, Stub(), topk=2, num_samples=S, device="cuda:0")
    ref = mpes_oracle.rank_mpes_from_samples(fs[:, :Q * C].reshape(S, Q, C).astype(np.float64),
                                             fs[:, Q * C:].astype(np.float64), topk=2)
    np.testing.assert_allclose(mi.numpy(), ref, rtol=1e-4, atol=1e-9)


def test_rank_pes_explicit_table_and_random_subset(pbo):
    """Explicit permutation tables (the reference's >1000-permutation path ships a random table)."""
    rng = np.random.default_rng(4)
    S, Q, C, M, k = 200, 11, 6, 4, 3
    fs = (rng.standard_normal((S, 3)) @ rng.standard_normal((3, Q * C + M))).astype(np.float32)
    fchi = torch.from_numpy(fs[:, :Q * C].reshape(S, Q, C).copy()).cuda()
    fstar = torch.from_numpy(fs[:, Q * C:].copy()).cuda()
    from importlib import import_module
    common = import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
    perms = mpes_oracle.all_permutations_k_in_n(C, k)                               # 120 rows: generic kernel
    sub = perms[rng.integers(0, 120, size=37)]                                      # duplicates allowed
    for table in (perms, sub):
        mi = common.mpes_from_device_samples(fchi, fstar, k, 0, table).cpu().numpy()
        ref = mpes_oracle.rank_mpes_streaming(fs[:, :Q * C].reshape(S, Q, C).astype(np.float64),
                                              np.argmax(fs[:, Q * C:], axis=1), M, table)
        np.testing.assert_allclose(mi, ref, rtol=1e-4, atol=1e-9)
    # all-permutation call without a table for a (C, k) that has no register-resident instantiation
    mi_all = common.mpes_from_device_samples(fchi, fstar, k, 0, None).cpu().numpy()
    ref_all = mpes_oracle.rank_mpes_from_samples(fs[:, :Q * C].reshape(S, Q, C).astype(np.float64),
                                                 fs[:, Q * C:].astype(np.float64), topk=k)
    np.testing.assert_allclose(mi_all, ref_all, rtol=1e-4, atol=1e-9)


def test_rank_pes_properties(pbo):
    """MI >= 0; MI == 0 when the query samples are independent of which maximiser wins; topk=1 == pes.I up to eps."""
    rng = np.random.default_rng(11)
    S, Q, C, M = 512, 40, 4, 6
    fchi = rng.standard_normal((S, Q, C)).astype(np.float32)
    fstar = rng.standard_normal((S, M)).astype(np.float32)
    fs = np.concatenate([fchi.reshape(S, Q * C), fstar], axis=1)
    mi = pbo.acquisitions.rank_pes.I_batch(np.zeros((Q, C, 1)), np.zeros((M, 1)), None, topk=2, num_samples=S,
                                           f_samples=fs, device="cuda:0").numpy()
    assert np.all(mi > -1e-9)
    fstar_const = np.zeros((S, M), dtype=np.float32)
    fstar_const[:, 2] = 1.0                                  # one maximiser always wins -> no information
    fs0 = np.concatenate([fchi.reshape(S, Q * C), fstar_const], axis=1)
    mi0 = pbo.acquisitions.rank_pes.I_batch(np.zeros((Q, C, 1)), np.zeros((M, 1)), None, topk=2, num_samples=S,
                                            f_samples=fs0, device="cuda:0").numpy()
    assert np.max(np.abs(mi0)) < 1e-6
    mi1 = pbo.acquisitions.rank_pes.I_batch(np.zeros((Q, C, 1)), np.zeros((M, 1)), None, topk=1, num_samples=S,
                                            f_samples=fs, device="cuda:0").numpy()
    for q in range(0, Q, 7):
        p = float(pbo.acquisitions.pes.I(np.zeros((C, 1)), np.zeros((M, 1)), None, num_samples=S, device="cuda:0",
                                         f_samples=np.concatenate([fstar, fchi[:, q, :]], axis=1)))
        assert abs(p - mi1[q]) < 1e-5


@pytest.mark.parametrize("tag", ["c3", "c5", "c2_t0", "gp_small", "gp_tiny"])
def test_indiff_pes_matches_reference_golden(pbo, golden, tag):
    """indiff_pes.I_batch (tkbo_mpes_indiff) on the samples the reference consumed -> the reference's MI."""
    g = golden("indiff_pes.npz")
    S, Q, C, M = (int(v) for v in g[f"{tag}_meta"])
    thr = float(g[f"{tag}_threshold"])
    fs = g[f"{tag}_fsample_all"]
    rng = np.random.default_rng(0)
    chi, x_star = rng.uniform(0, 1, (Q, C, 4)), rng.uniform(0, 1, (M, 4))          # only their shapes matter with f_samples
    mi = pbo.acquisitions.indiff_pes.I_batch(chi, x_star, None, num_samples=S, indifference_threshold=thr,
                                             f_samples=fs, device="cuda:0")
    ref = g[f"{tag}_mi"]
    assert mi.dtype == torch.float64 and tuple(mi.shape) == (Q,)
    np.testing.assert_allclose(mi.numpy(), ref, rtol=VAL_RTOL, atol=1e-9)
    if f"{tag}_loglik" in g.files:
        ll = pbo.acquisitions.indiff_pes.get_log_likelihood(fs[:, :Q * C].reshape(S, Q, C), S, thr, device="cuda:0")
        rows = C + 1 if thr > 0 else C        # threshold 0: P(none) is pure rounding noise in the reference as well
        np.testing.assert_allclose(ll.numpy()[:rows], g[f"{tag}_loglik"][:rows], rtol=1e-9, atol=1e-10)


def test_indiff_pes_large_batch_matches_oracle(pbo):
    rng = np.random.default_rng(21)
    S, Q, C, M = 512, 3000, 4, 10
    basis = rng.standard_normal((Q * C + M, 6))
    fs = (rng.standard_normal((S, 6)) @ basis.T + 0.3 * rng.standard_normal((S, Q * C + M))).astype(np.float32)
    mi = pbo.acquisitions.indiff_pes.I_batch(np.zeros((Q, C, 2)), np.zeros((M, 2)), None, num_samples=S,
                                             indifference_threshold=0.2, f_samples=fs, device="cuda:0").numpy()
    ref = mpes_oracle.indiff_mpes_from_samples(fs[:, :Q * C].reshape(S, Q, C).astype(np.float64), fs[:, Q * C:].astype(np.float64), 0.2)
    np.testing.assert_allclose(mi, ref, rtol=VAL_RTOL, atol=1e-9)
    assert mi.min() > -1e-9


@pytest.mark.parametrize("tag", ["c1", "c3", "small"])
def test_pes_I_matches_reference_golden(pbo, golden, tag):
    """pes.I / pes.I_batch executed from the reference source (golden) vs K2 in PES mode, c1 sizes S=1000, M=20, C=2."""
    g = golden("pes.npz")
    S, M, C, Qn = [int(v) for v in g[f"{tag}_meta"]]
    smp = g[f"{tag}_samples"]
    got = np.array([float(pbo.acquisitions.pes.I(np.zeros((C, 1)), np.zeros((M, 1)), None, num_samples=S,
                                                 f_samples=smp[q], device="cuda:0")) for q in range(Qn)])
    np.testing.assert_allclose(got, g[f"{tag}_I"], rtol=1e-4, atol=1e-9)


def test_pes_I_batch_reference_loop_semantics(pbo, golden):
    """I_batch with a model that only has predict_f_samples: one draw per query, like pes.py:132."""
    g = golden("pes.npz")
    S, M, C, Qn = [int(v) for v in g["c1_meta"]]
    smp = g["c1_samples"]

    class Stub:
        calls = 0

        def predict_f_samples(self, x, n):
            assert x.shape == (M + C, 1) and n == S
            out = torch.from_numpy(smp[Stub.calls].astype(np.float64)[:, :, None])
            Stub.calls += 1
            return out

    vals = pbo.acquisitions.pes.I_batch(np.zeros((Qn, C, 1)), np.zeros((M, 1)), Stub(), device="cuda:0")
    assert isinstance(vals, np.ndarray) and vals.shape == (Qn,)
    np.testing.assert_allclose(vals, g["c1_I"], rtol=1e-4, atol=1e-9)


def test_dts_sample_f_and_copeland_match_reference_golden(pbo, golden):
    """dts.sample_f on the basis the reference drew (W, b, z from the golden file) and the soft-Copeland winner."""
    g = golden("dts.npz")
    ff = pbo.fourier_features
    n = g["Xq"].shape[0]
    phi_y = ff.fourier_features(g["Xq"][None], g["W"], g["b"], device="cuda:0")
    theta = ff.sample_theta_variational(phi_y, g["q_mu"], g["q_sqrt"], z=g["z"])            # (1, D, 1)
    basis = ff.SharedBasis(g["W"][0], g["b"][0, :, 0], theta[0], device="cuda:0")
    f = basis.eval(g["combs"]).reshape(-1, 1)
    ref = g["f_vals"]
    assert np.max(np.abs(f.cpu().numpy() - ref)) < VAL_RTOL * np.abs(ref).max()
    winner = pbo.acquisitions.dts.soft_copeland_maximizer(f, g["grid"])
    np.testing.assert_array_equal(winner, g["winner"])
    # CPU-tensor input path and the scores themselves
    w2, scores = pbo.acquisitions.dts.soft_copeland_maximizer(ref, g["grid"], device="cuda:0", return_scores=True)
    np.testing.assert_array_equal(w2, g["winner"])
    np.testing.assert_allclose(scores.numpy(), mpes_oracle.soft_copeland_scores(ref), rtol=1e-5)
    assert n == 9


def test_dts_sample_f_signature_path(pbo):
    """sample_f(m, query_points, combs, D) with a Product kernel: shape, dtype, finite, Copeland winner on the grid."""
    dts = pbo.acquisitions.dts
    mdl = pbo.model.synthetic_model(9, 4, 0.0, 1.0, [0.25, 0.4, 0.25, 0.4], seed=5, product=True)
    grid = dts.uniform_grid(2, 10, 0, 1)
    combs = dts.combinations(grid)
    f = dts.sample_f(mdl, mdl.Z, combs, 64, device="cuda:0")
    assert tuple(f.shape) == (10000, 1) and f.dtype == torch.float64 and bool(torch.isfinite(f).all())
    assert dts.soft_copeland_maximizer(f, grid, device="cuda:0").shape == (2,)


@pytest.mark.parametrize("n,dobj,D,S", [(300, 2, 256, 40), (128, 1, 128, 300), (257, 3, 192, 8)])
def test_duel_copeland_fused_matches_oracle_and_unfused_path(pbo, n, dobj, D, S):
    """Fused duel generation + soft-Copeland epilogue (dts.py:31-45, 71-97 for S samples at once) against the fp64
    oracle on the explicit n^2 duels, and against the unfused route (store-F kernel + tkbo_soft_copeland)."""
    rng = np.random.default_rng(n + D + S)
    Xb = rng.uniform(0.0, 1.0, size=(n, dobj)).astype(np.float32)
    W = rng.standard_normal((D, 2 * dobj)) / 0.4
    b = rng.uniform(0, 2 * np.pi, D)
    Theta = rng.standard_normal((D, S)) * rng.uniform(0.5, 3.0, size=(1, S))
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    score, win = basis.duel_copeland(Xb)
    combs = pbo.acquisitions.dts.combinations(Xb.astype(np.float64))
    F = rff_oracle.rff_values_shared(combs, W, b, Theta)                       # (n^2, S)
    ref = np.stack([mpes_oracle.soft_copeland_scores(F[:, s]) for s in range(S)])   # (S, n)
    np.testing.assert_allclose(score.cpu().numpy(), ref, rtol=2e-5, atol=1e-6)
    ref_win = ref.argmax(axis=1)
    got = win.cpu().numpy()
    top = np.sort(ref, axis=1)
    clear = (top[:, -1] - top[:, -2]) > 1e-5                                  # winners with a resolvable margin
    np.testing.assert_array_equal(got[clear], ref_win[clear])
    assert np.all(ref[np.arange(S), got] >= top[:, -1] - 1e-5)
    # deterministic fixed-point sums: bitwise repeatable
    score2, win2 = basis.duel_copeland(Xb)
    assert torch.equal(score, score2) and torch.equal(win, win2)
    # unfused route on the first sample
    f0 = basis.eval(combs.astype(np.float32))[0].reshape(-1, 1)
    _, sc0 = pbo.acquisitions.dts.soft_copeland_maximizer(f0, Xb, return_scores=True)
    np.testing.assert_allclose(score[0].cpu().numpy(), sc0.numpy(), rtol=2e-5, atol=1e-6)


def test_duel_copeland_shards_add_up_exactly(pbo):
    """j-shards of the duels (the multi-GPU decomposition) summed as integers == one call, bit for bit."""
    rng = np.random.default_rng(77)
    n, dobj, D, S = 200, 2, 128, 24
    Xb = rng.uniform(0.0, 1.0, size=(n, dobj)).astype(np.float32)
    basis = pbo.fourier_features.SharedBasis(rng.standard_normal((D, 2 * dobj)) / 0.5, rng.uniform(0, 2 * np.pi, D),
                                             rng.standard_normal((D, S)), device="cuda:0")
    score, win = basis.duel_copeland(Xb)
    fix = None
    for r in range(3):
        lo, hi = pbo.distributed.duel_shard(n, r, 3)
        part = basis.duel_copeland_partial(Xb, lo, hi)
        fix = part if fix is None else fix + part
    s2, w2 = basis.copeland_finish(fix)
    assert torch.equal(score, s2) and torch.equal(win, w2)
    s3, w3 = pbo.distributed.sharded_duel_copeland(basis, Xb)            # no process group: single shard
    assert torch.equal(score, s3) and torch.equal(win, w3)


def test_dts_thompson_copeland_winners_signature(pbo):
    dts = pbo.acquisitions.dts
    mdl = pbo.model.synthetic_model(9, 4, 0.0, 1.0, [0.25, 0.4, 0.25, 0.4], seed=5, product=True)
    grid = dts.uniform_grid(2, 12, 0, 1)
    winners, scores = dts.thompson_copeland_winners(mdl, mdl.Z, grid, D=64, count=5, device="cuda:0", return_scores=True)
    assert winners.shape == (5, 2) and tuple(scores.shape) == (5, 144)
    assert bool(torch.isfinite(scores).all()) and float(scores.min()) > 0.0 and float(scores.max()) < 1.0


def test_rff_joint_sampling_feeds_mpes(pbo):
    """model.predict_f_samples_device (RFF draws through the store-F epilogue) -> K2, nothing leaves the GPU."""
    mdl = pbo.model.synthetic_model(16, 6, 0.0, 1.0, 0.4, seed=8)
    rng = np.random.default_rng(0)
    chi = rng.uniform(size=(300, 5, 6))
    xs = rng.uniform(size=(12, 6))
    mi = pbo.acquisitions.rank_pes.I_batch(chi, xs, mdl, topk=3, num_samples=256, rff_features=256, device="cuda:0")
    assert tuple(mi.shape) == (300,) and bool(torch.isfinite(mi).all()) and float(mi.min()) > -1e-9
    vals = pbo.acquisitions.pes.I_batch(chi[:50, :2], xs, mdl, device="cuda:0")
    assert vals.shape == (50,) and np.isfinite(vals).all()


def test_large_shape_properties(pbo):
    """Size-independent properties at a BASELINE-sized tile count (c2: N = 2^20, D = 1024, S = 64): the argmax
    is invariant to candidate order up to the permutation, the fused max equals the max of the store-F output on
    a slice containing the winner, and sharding leaves it unchanged."""
    rng = np.random.default_rng(1)
    N, D, d, S = 1 << 20, 1024, 2, 64
    W = rng.standard_normal((D, d)) / 0.6
    b = rng.uniform(0, 2 * np.pi, D)
    Theta = rng.standard_normal((D, S))
    X = torch.rand((N, d), device="cuda:0") * 3.0 - 1.5
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0")
    val, idx = basis.argmax(X)
    perm = torch.randperm(N, device="cuda:0")
    val_p, idx_p = basis.argmax(X[perm].contiguous())
    assert torch.equal(val, val_p)
    # at N = 2^20 distinct candidates can tie bit-for-bit in fp32, so the winner may change with the order;
    # what must hold is that every returned index attains the returned maximum
    rows = basis.eval(X[idx].contiguous())                     # (S, S): sample s at its own winner
    assert torch.equal(rows.diagonal(), val)
    rows_p = basis.eval(X[perm[idx_p]].contiguous())
    assert torch.equal(rows_p.diagonal(), val)
    assert int((perm[idx_p] != idx).sum()) <= 8               # exact cross-candidate ties are rare
    half = N // 2
    v0, i0 = basis.argmax(X[:half])
    v1, i1 = basis.argmax(X[half:], idx_offset=half)
    mv, mi = pbo.distributed.merge_pairs(torch.stack([v0, v1]).cpu(), torch.stack([i0, i1]).cpu())
    assert torch.equal(mv, val.cpu()) and torch.equal(mi, idx.cpu())
    # oracle spot check on the winners' neighbourhood only (cheap): fp64 value at the returned index
    xw = X[idx].cpu().numpy().astype(np.float64)
    f64 = rff_oracle.rff_values_shared(xw, W, b, Theta)
    assert np.max(np.abs(np.diag(f64) - val.cpu().numpy()) / np.abs(np.diag(f64))) < VAL_RTOL


def test_large_shape_properties_tensor_bound_form(pbo):
    """The same properties at a c3-like shape (d = 6, D = 4096, S = 256: one 256-sample chunk, e4m3 correction form,
    separate (W | b) ring, slab-sharing producers), plus agreement of the fast and fp32 contraction forms."""
    rng = np.random.default_rng(2)
    N, D, d, S = 1 << 19, 4096, 6, 256
    W = rng.standard_normal((D, d)) / 0.2
    b = rng.uniform(0, 2 * np.pi, D)
    Theta = rng.standard_normal((D, S))
    X = torch.rand((N, d), device="cuda:0")
    basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision="fast")
    cfg = basis.config(N)
    assert cfg["e4m3_corrections"] == 1 and cfg["S_chunk"] == 256
    assert (cfg["cta_pairs"] == 1 and cfg["stages"] >= 5) or (cfg["w_stages"] >= 2 and cfg["stages"] == 3)
    val, idx = basis.argmax(X)
    perm = torch.randperm(N, device="cuda:0")
    val_p, idx_p = basis.argmax(X[perm].contiguous())
    assert torch.equal(val, val_p)
    rows = basis.eval(X[idx].contiguous())
    assert torch.equal(rows.diagonal(), val)
    third = (N // 3) // 128 * 128
    parts = [basis.argmax(X[lo:hi], idx_offset=lo) for lo, hi in ((0, third), (third, 2 * third), (2 * third, N))]
    mv, mi = pbo.distributed.merge_pairs(torch.stack([p[0] for p in parts]).cpu(), torch.stack([p[1] for p in parts]).cpu())
    assert torch.equal(mv, val.cpu()) and torch.equal(mi, idx.cpu())
    xw = X[idx].cpu().numpy().astype(np.float64)
    f64 = np.diag(rff_oracle.rff_values_shared(xw, W, b, Theta))
    assert np.max(np.abs(f64 - val.cpu().numpy()) / np.abs(f64)) < VAL_RTOL
    precise = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision="fp32")
    val32, idx32 = precise.argmax(X)
    assert np.max(np.abs(f64 - val32.cpu().numpy()[np.arange(S)]) / np.abs(f64)) < 1e-4
    rel = (val - val32).abs() / val32.abs()
    assert float(rel.max()) < VAL_RTOL
    # the two forms may pick different winners only where the top-2 gap is below the tolerance
    diff = (idx != idx32).nonzero().flatten()
    if diff.numel():
        f_at_fast = precise.eval(X[idx[diff]].contiguous())[diff, torch.arange(diff.numel(), device="cuda:0")]
        assert float(((val32[diff] - f_at_fast) / val32[diff].abs()).max()) < VAL_RTOL


def test_c3_shape_regression_weights_both_forms(pbo):
    """c3 shape with the weights the bench uses (the ff.py:56-87 regression on a synthetic fitted posterior, not Gaussian
    Theta): every value on a 2^15-candidate sample against the fp64 oracle in BOTH contraction forms, the tie count, and
    the form precision="auto" chooses from the training interpolation (SharedBasis.sample passes max|q| as f_scale)."""
    rng = np.random.default_rng(5)
    N, D, d, S, n_train = 1 << 15, 4096, 6, 256, 41
    X, W, b, Theta = _basis_inputs(rng, N, D, d, S, 0.0, 1.0, 0.2, n_train=n_train)
    Fref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)
    scale = np.abs(Fref).max(axis=0)
    errs, ties = {}, {}
    for precision in ("fp32", "fast"):
        basis = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision=precision)
        F = basis.eval(X).cpu().numpy().T.astype(np.float64)
        errs[precision] = float(np.max(np.abs(F - Fref) / scale))
        val, idx = basis.argmax(X)
        ties[precision] = _check_argmax(val, idx, X, W, b, Theta)
    norm_ratio = float((np.linalg.norm(Theta * np.sqrt(2.0 / D), axis=0) / scale).max())
    print(f"c3-shape regression weights: max|dv|/max|f| fp32 {errs['fp32']:.2e} fast {errs['fast']:.2e}, ties {ties}, "
          f"||theta'|| / max|f| <= {norm_ratio:.2f}")
    # at D = 4096 the error of BOTH forms is set by the tensor core's truncating fp32 accumulation (a bias towards zero of
    # ~1.2e-8 max|f| per accumulator update, tools/acc_bias_probe.py): 3 D / 16 updates in the fp32 form, 2 D / 16 in the fast
    assert errs["fp32"] < 1e-5 + 2.5e-8 * 3 * D / 16 and errs["fast"] < 1e-5 + 2.5e-8 * 2 * D / 16 + 1e-5
    assert ties["fp32"] <= 2 and ties["fast"] <= 8
    # auto: the regression interpolates q on the training inputs, ||theta'|| / max|q| ~ 2-3 here -> fp32
    q_scale = scale                                                       # max|f| >= max|q|; be generous to the fast form
    auto = pbo.fourier_features.SharedBasis(W, b, Theta, device="cuda:0", precision="auto", f_scale=q_scale)
    assert auto.precision == ("fast" if norm_ratio <= 0.3 else "fp32")


def test_mpes_query_shards_equal_full(pbo):
    """K2 is independent per query set given the (global) samples: the MI of a query shard computed alone is bitwise
    the MI of those queries in the full batch, for ragged shard boundaries too, and the sharded best query is np.argmax."""
    gen = torch.Generator(device="cuda:0").manual_seed(11)
    S, Q, C, k, M = 512, 3000, 5, 3, 12
    fchi = torch.randn((S, Q, C), device="cuda:0", generator=gen)
    fstar = torch.randn((S, M), device="cuda:0", generator=gen)
    common = pbo.acquisitions._common
    nat = pbo._native
    full = common.mpes_from_device_samples(fchi, fstar, k, nat.MPES_RANK)
    dist_ = pbo.distributed
    parts, bests = [], []
    for r in range(3):
        lo, hi = dist_.query_shard(Q, r, 3)
        mi, best = dist_.sharded_mpes(fchi[:, lo:hi, :].contiguous(), fstar, k, nat.MPES_RANK, lo)
        parts.append(mi)
        bests.append(best)
    assert torch.equal(torch.cat(parts), full)
    top = max(bests, key=lambda b: (b[0], -b[1]))
    assert top[1] == int(torch.argmax(full)) and top[0] == float(full.max())
    mi0, best0 = dist_.sharded_mpes(fchi[:, :0, :].contiguous(), fstar, k, nat.MPES_RANK, 0)
    assert mi0.numel() == 0 and best0 == (-float("inf"), -1)


def test_bench_gpu_arm_contract():
    """`bench.py` (GPU arm) on the small workload: one JSON line with the contract's keys, a roofline and an e2e object
    whose result equals the device-resident one, kernels launched through libtkbo.so, same `config` dict as the reference
    arm prints for that workload."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--workload", "c2", "--steps", "5", "--warmup", "3",
                        "--no-blocks", "--no-cpu"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert key in j, key
    assert j["n_gpus"] == 1 and j["steps"] == 5 and j["gpu_launches"] in (5, 15) and j["value"] > 1e10
    assert j["config"] == {"workload": j["config"]["workload"], "N_total": 1 << 20, "d": 2, "D": 1024, "S": 64}
    assert j["e2e"]["equals_device_resident_result"] is True and j["e2e"]["h2d_bytes_per_step"] > 8e6
    rf = j["roofline"]
    assert rf["bound"] == "tensor" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["form"] == "fp32"
