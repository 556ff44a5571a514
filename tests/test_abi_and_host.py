"""CPU-side checks: the C-ABI library loads and exports every symbol include/tkbo.h declares (no compute calls
without a GPU), the product fails loudly without CUDA, host-side helpers agree with the oracle."""
import ctypes
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import mpes_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pbo):
    nat = pbo._native
    declared = nat.declared_symbols()
    assert len(declared) >= 18 and "tkbo_rff_argmax" in declared and "tkbo_mpes_topk" in declared
    lib = ctypes.CDLL(nat.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), f"libtkbo.so does not export {name}"
    # and the binding table covers exactly the header
    assert sorted(nat._SIGNATURES) == declared
    assert nat.lib().tkbo_version() == 100


def test_library_is_sm100a_with_tcgen05_and_tma(pbo):
    """The shipped binary really contains the Blackwell paths (UTCHMMA = tcgen05.mma, UTMALDG = TMA, LDTM/STTM)."""
    out = subprocess.run(["cuobjdump", "-sass", pbo._native.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    sass = out.stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "UTMALDG", "UTCBAR.2CTA.MULTICAST", "LDTM", "STTM", "CREDUX", "LDGSTS"):
        assert mnemonic in sass, mnemonic


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(pbo):
    with pytest.raises(pbo.TkboError):
        pbo.fourier_features.SharedBasis(np.zeros((64, 2)), np.zeros(64), np.zeros((64, 4)))
    with pytest.raises(pbo.TkboError):
        pbo.acquisitions.rank_pes.I_batch(np.zeros((2, 3, 1)), np.zeros((2, 1)), None, f_samples=np.zeros((8, 8)))
    out = ctypes.c_void_p()
    rc = pbo._native.lib().tkbo_create(ctypes.byref(out), 0)
    assert rc != 0 and "no CUDA device" in pbo._native.last_error()


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no file of the product package may reference it."""
    pkg_dir = os.path.join(ROOT, "top-k-ranking-bayesian-optimization_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("CPU oracle", ""), f"{f} mentions the oracle"


def test_import_by_directory_name_and_alias():
    a = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    import tkbo_b200
    assert tkbo_b200 is a
    for name in ("fourier_features", "sample_fourier_features", "sample_theta_variational", "sample_features_weights",
                 "sample_maximizers"):
        assert callable(getattr(a.fourier_features, name))
    assert callable(a.acquisitions.pes.I) and callable(a.acquisitions.pes.I_batch)
    assert callable(a.acquisitions.rank_pes.I_batch) and callable(a.acquisitions.dts.sample_f)
    assert callable(a.acquisitions.dts.soft_copeland_maximizer)


def test_signatures_match_reference_defaults(pbo):
    import inspect
    sig = inspect.signature(pbo.fourier_features.sample_maximizers)
    assert list(sig.parameters)[:8] == ["X", "count", "n_init", "D", "model", "min_val", "max_val", "num_steps"]
    assert sig.parameters["num_steps"].default == 3000                          # ff.py:110
    assert inspect.signature(pbo.fourier_features.sample_fourier_features).parameters["D"].default == 100   # ff.py:24
    sig = inspect.signature(pbo.acquisitions.rank_pes.I_batch)
    assert list(sig.parameters)[:6] == ["chi", "x_star", "model", "topk", "num_samples", "indifference_threshold"]
    assert sig.parameters["num_samples"].default == 10000 and sig.parameters["topk"].default is None   # rank_pes.py:10
    assert inspect.signature(pbo.acquisitions.pes.I).parameters["num_samples"].default == 1000        # pes.py:92
    assert inspect.signature(pbo.acquisitions.dts.sample_f).parameters["D"].default == 1000           # dts.py:71


def test_permutation_helpers_match_oracle(pbo, golden):
    rp = pbo.acquisitions.rank_pes
    g = golden("rank_pes.npz")
    np.testing.assert_array_equal(rp.get_all_permutation_k_in_n(5, 3), g["perms_5_3"])
    np.testing.assert_array_equal(rp.get_all_permutation_k_in_n(4, 2), g["perms_4_2"])
    np.testing.assert_array_equal(rp.precompute_n_permutations(5, 3), g["nperm_5_3"])
    np.random.seed(0)
    r = rp.get_rand_permutation_k_in_n(6, 4, 50)
    assert r.shape == (50, 4) and all(len(set(row)) == 4 for row in r.tolist())
    for n, k in [(3, 1), (4, 4), (6, 2)]:
        np.testing.assert_array_equal(rp.get_all_permutation_k_in_n(n, k), mpes_oracle.all_permutations_k_in_n(n, k))


def test_dts_grid_helpers(pbo, golden):
    g = golden("dts.npz")
    dts = pbo.acquisitions.dts
    np.testing.assert_array_equal(dts.uniform_grid(2, 12, 0, 1), g["grid"])
    np.testing.assert_array_equal(dts.combinations(g["grid"]), g["combs"])


def test_kernel_lengthscales_product_and_rbf(pbo):
    m = pbo.model
    assert m.kernel_lengthscales(m.RBF([0.2, 0.3]), 2).tolist() == [0.2, 0.3]
    assert m.kernel_lengthscales(m.RBF(0.2), 3).tolist() == [0.2, 0.2, 0.2]
    prod = m.Product([m.RBF(0.25), m.RBF(0.4)])
    assert m.is_product_kernel(prod) and not m.is_product_kernel(m.RBF(1.0))
    assert m.kernel_lengthscales(prod, 4).tolist() == [0.25, 0.4, 0.25, 0.4]      # ff.py:33-40


def test_exact_predict_f_samples_shape_and_moments(pbo):
    mdl = pbo.model.synthetic_model(12, 2, -1.5, 1.5, 0.6, seed=3)
    X = np.random.default_rng(0).uniform(-1.5, 1.5, (7, 2))
    s = mdl.predict_f_samples(X, 4000)
    assert tuple(s.shape) == (4000, 7, 1) and s.numpy().dtype == np.float64
    mu, cov = mdl.predict_f(X, full_cov=True)
    assert np.allclose(s.numpy()[:, :, 0].mean(0), mu[:, 0], atol=5 * np.sqrt(np.diag(cov).max() / 4000) + 1e-3)


def test_shard_ranges_cover_and_align(pbo):
    sr = pbo.distributed.shard_range
    for N, G in [(1000, 3), (1 << 20, 8), (129, 4), (5, 2)]:
        parts = [sr(N, r, G) for r in range(G)]
        assert parts[0][0] == 0 and parts[-1][1] == N
        for (a, b), (c, d) in zip(parts, parts[1:]):
            assert b == c and (a % 128 == 0 or a == b)
        tiles = [(b - a + 127) // 128 for a, b in parts]
        assert max(tiles) - min(tiles) <= 1


def test_merge_pairs_lowest_index_on_ties(pbo):
    vals = torch.tensor([[1.0, 5.0, 2.0], [1.0, 4.0, 3.0], [0.5, 5.0, 3.0]])
    idxs = torch.tensor([[10, 11, 12], [3, 4, 5], [7, 2, -1]])
    v, i = pbo.distributed.merge_pairs(vals, idxs)
    assert v.tolist() == [1.0, 5.0, 3.0] and i.tolist() == [3, 2, 5]


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    F = rng.standard_normal((1000, 16)).astype(np.float32)          # full (N, S) sample matrix, same on all ranks
    F[500, 3] = F[100, 3] = F[:, 3].max() + 1.0                      # exact tie across shards -> lowest index wins
    lo, hi = pkg.distributed.shard_range(1000, rank, world)
    loc = torch.from_numpy(F[lo:hi])
    val, idx = loc.max(dim=0)
    # torch.max returns an arbitrary tied index; recompute the lowest like the kernel does
    idx = torch.from_numpy(np.argmax(F[lo:hi], axis=0)) + lo
    v, i = pkg.distributed.allgather_argmax(val, idx)
    q.put((rank, v.numpy(), i.numpy()))
    dist.destroy_process_group()


def test_allgather_merge_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    F = rng.standard_normal((1000, 16)).astype(np.float32)
    F[500, 3] = F[100, 3] = F[:, 3].max() + 1.0
    for _, v, i in res:
        np.testing.assert_array_equal(i, np.argmax(F, axis=0))
        np.testing.assert_array_equal(v, F.max(axis=0))
    assert res[0][2][3] == 100


def _gloo_key_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    F = rng.standard_normal((900, 12)).astype(np.float32)
    F[700, 5] = F[40, 5] = F[:, 5].max() + 1.0                       # exact tie across shards -> lowest index wins
    F[:, 7] = -np.abs(F[:, 7])                                       # an all-negative sample (sign handling of the key)
    lo, hi = pkg.distributed.shard_range(900, rank, world)
    val = torch.from_numpy(F[lo:hi].max(axis=0))
    idx = torch.from_numpy(np.argmax(F[lo:hi], axis=0).astype(np.int64)) + lo
    key = pkg.distributed.pack_argmax_key(val, idx)
    dist.all_reduce(key, op=dist.ReduceOp.MAX)                       # what allreduce_argmax does with NCCL
    v, i = pkg.distributed.unpack_argmax_key(key)
    # duel shards: integer partial sums add up exactly
    jlo, jhi = pkg.distributed.duel_shard(101, rank, world)
    part = torch.arange(jlo, jhi, dtype=torch.int64).sum().reshape(1)
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    q.put((rank, v.numpy(), i.numpy(), int(part)))
    dist.destroy_process_group()


def test_key_merge_and_duel_shards_world2_gloo():
    """The one-collective merges of the multi-GPU paths (int64 all_reduce MAX of packed keys; SUM of integer duel
    sums) on two gloo ranks, with the torch restatement of the device-side key."""
    import torch.multiprocessing as mp
    pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    v0 = torch.tensor([3.5, -2.25, 0.0, -0.0, float("-inf"), 1e-30], dtype=torch.float32)
    i0 = torch.tensor([7, 123456789, 0, 5, 2, -1])
    v1, i1 = pkg.distributed.unpack_argmax_key(pkg.distributed.pack_argmax_key(v0, i0))
    assert torch.equal(v1[:5], v0[:5]) and torch.equal(i1[:5], i0[:5]) and int(i1[5]) == -1 and bool(torch.isnan(v1[5]))
    ks = pkg.distributed.pack_argmax_key(torch.tensor([1.0, 1.0, -1.0, 2.0]), torch.tensor([9, 4, 0, 100]))
    assert ks[1] > ks[0] > ks[2] and ks[3] > ks[1]                   # value first, then the lower index
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_key_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(1)
    F = rng.standard_normal((900, 12)).astype(np.float32)
    F[700, 5] = F[40, 5] = F[:, 5].max() + 1.0
    F[:, 7] = -np.abs(F[:, 7])
    for _, v, i, part in res:
        np.testing.assert_array_equal(i, np.argmax(F, axis=0))
        np.testing.assert_array_equal(v, F.max(axis=0))
        assert part == sum(range(101))
    assert res[0][2][5] == 40


def test_query_set_generators_reproduce_reference_under_seed(golden):
    """pes.sample_inputs / sample_inputs_discrete: same np.random stream as pes.py:135-189 -> identical query sets."""
    import importlib
    pes = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions.pes")
    g = golden("pes_inputs.npz")
    np.random.seed(1234)
    a = pes.sample_inputs(g["cur"], 6, 3, -1.0, 2.0)
    np.random.seed(4321)
    b = pes.sample_inputs_discrete(g["cur"], g["data"], 5, 4)
    assert tuple(a.shape) == (24, 3, 3) and tuple(b.shape) == (20, 4, 3)
    np.testing.assert_array_equal(a.numpy(), g["uniform"])
    np.testing.assert_array_equal(b.numpy(), g["discrete"])


def test_dts_gauss_hermite_helpers_match_numerical_integration():
    """dts.expected_integral / variance_integral / variance_logistic_f (dts.py:48-68) against scipy quadrature."""
    from scipy import integrate
    dts = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions.dts")
    means = np.array([[-2.0], [0.0], [0.7], [3.0]])
    variances = np.array([[[0.01]], [[1.0]], [[4.0]], [[0.5]]])
    sig = lambda t: 1.0 / (1.0 + np.exp(-t))   # noqa: E731
    for fn, g in ((dts.expected_integral, sig), (dts.variance_integral, lambda t: sig(t) ** 2)):
        got = fn(means, variances).numpy()[:, 0]
        ref = [integrate.quad(lambda t, m=m, v=v: g(t) * np.exp(-0.5 * (t - m) ** 2 / v) / np.sqrt(2 * np.pi * v), m - 12 * np.sqrt(v),
                              m + 12 * np.sqrt(v))[0] for m, v in zip(means[:, 0], variances[:, 0, 0])]
        np.testing.assert_allclose(got, ref, rtol=2e-6, atol=1e-9)        # H = 50 nodes on a non-polynomial integrand

    class M:
        def predict_f(self, x):
            return means, variances[:, 0]
    v = dts.variance_logistic_f(M(), np.zeros((4, 2))).numpy()
    assert v.shape == (4, 1) and np.all(v > 0) and np.all(v < 0.25)


def test_pes_building_blocks_reassemble_to_I():
    """pes.samples_likelihood / log_p_x_star_cond_D / log_p_z_cond_D_x_star (pes.py:44-89) combined as pes.I does
    (pes.py:108-123) equal the oracle's restatement of I."""
    pes = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions.pes")
    rng = np.random.default_rng(5)
    S, M, C = 400, 6, 3
    samples = rng.standard_normal((S, 3)) @ rng.standard_normal((3, M + C)) + 0.3 * rng.standard_normal((S, M + C))
    arg = samples[:, :M].argmax(axis=1)
    log_px = pes.log_p_x_star_cond_D(M, arg)
    total = 0.0
    for i in range(C):
        z = (i, None)
        log_pz_x = pes.log_p_z_cond_D_x_star(M, samples, arg, z, log_px)
        log_pz = torch.logsumexp(log_px + log_pz_x, dim=0)                      # pes.py:115-118
        total += float((torch.exp(log_px) * torch.exp(log_pz_x) * (log_pz_x - log_pz)).sum())
    np.testing.assert_allclose(total, mpes_oracle.pes_I_from_samples(samples, M), rtol=1e-9)


def _gloo_query_worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    rng = np.random.default_rng(5)
    mi = rng.random(1000)                                             # the full MI vector, same on all ranks
    mi[777] = mi[40] = mi.max() + 0.5                                 # exact tie across shards -> lowest query index
    lo, hi = pkg.distributed.query_shard(1000, rank, world)
    loc = torch.from_numpy(mi[lo:hi])
    best = pkg.distributed.best_query(loc, lo)
    full = pkg.distributed.allgather_queries(loc, 1000)
    # a rank without queries (more ranks than 32-query tiles) contributes nothing
    lo2, hi2 = pkg.distributed.query_shard(20, rank, world)
    best2 = pkg.distributed.best_query(torch.from_numpy(mi[lo2:hi2]), lo2)
    q.put((rank, best, full.numpy(), (lo2, hi2), best2))
    dist.destroy_process_group()


def test_query_shards_best_and_gather_world2_gloo():
    """K2's multi-GPU side: contiguous query shards of whole 32-query tiles, the (best MI, lowest index)
    pair over all shards, and the gathered MI vector, on two gloo ranks."""
    import torch.multiprocessing as mp
    pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    for Q, world in ((1000, 2), (100000, 8), (31, 4), (33, 2)):
        ranges = [pkg.distributed.query_shard(Q, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == Q
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        assert all(lo % 32 == 0 or lo == Q for lo, _ in ranges)
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 32 + (Q % 32 != 0) * 32
    assert pkg.distributed.best_query(torch.tensor([0.1, 0.7, 0.7, 0.2]), 10) == (pytest.approx(0.7), 11)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_query_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    mi = rng.random(1000)
    mi[777] = mi[40] = mi.max() + 0.5
    for _, best, full, (lo2, hi2), best2 in res:
        assert best == (mi[40], 40)
        np.testing.assert_array_equal(full, mi)
        assert best2 == (mi[:20].max(), int(np.argmax(mi[:20])))
    assert sorted(r[3] for r in res) == [(0, 20), (20, 20)]           # the second rank's shard is empty


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver times beside the GPU arm) runs without a GPU and prints one
    JSON line with the contract's keys: same metric / unit / config as the GPU arm, impl = reference, a cpu_baseline
    describing the run, and an e2e object without transfers."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--budget-s", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["n_gpus"] == 1 and j["steps"] == 1
    assert j["metric"].startswith("RFF posterior-sample candidate*samples/s") and j["unit"] == "candidate*samples/s"
    assert j["higher_is_better"] is True and j["value"] > 0 and j["dtype"] == "f64"
    # the default workload is the largest single-GPU configuration (c3); `config` is exactly the GPU arm's dict
    assert j["config"] == {"workload": j["config"]["workload"], "N_total": 1 << 24, "d": 6, "D": 4096, "S": 256}
    assert j["config"]["workload"].startswith("c3") and j["scaling"] == "strong"
    cb = j["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == j["value"] and cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_two_pass_bound_covers_the_correction_terms_in_simulation():
    """The two-pass argmax (csrc/rff.cu: launch_argmax) excludes rows whose hi*hi-only value lies more than 2 eps_s below the
    largest one, with eps_s = A * sum_j |hi(theta'_sj)| * (2^-9 + Dpad/4 * 2^-23) (prep_basis_kernel).  NumPy / torch restatement
    of both operand splits - fp16 hi/lo (fp32 form) and fp16 hi + e4m3 corrections on u = 4096 cos (fast form) - on random
    data: the dropped terms never exceed eps_s, with the margin DESIGN section 3 states."""
    import torch
    rng = np.random.default_rng(5)
    D, S, R = 512, 8, 400
    theta = rng.standard_normal((D, S)) * rng.uniform(0.2, 5.0, size=(1, S)) * np.sqrt(2.0 / D)
    phi = np.cos(rng.uniform(-40.0, 40.0, size=(R, D)))

    def f16(a):
        return a.astype(np.float16).astype(np.float64)

    def e4m3(a):
        return torch.from_numpy(np.asarray(a, dtype=np.float32)).to(torch.float8_e4m3fn).to(torch.float64).numpy()

    worst = {"fp32": 0.0, "fast": 0.0}
    for s in range(S):
        amax = np.abs(theta[:, s]).max()
        # fp32 form: max |v| in [2^9, 2^10); fast form: [2^7, 2^8)
        for form, top, amp in (("fp32", 9, 1.0), ("fast", 7, 4096.0)):
            v = theta[:, s] * 2.0 ** (top - np.floor(np.log2(amax)))
            vh = f16(v)
            eps = amp * np.abs(vh).sum() * (2.0 ** -9 + (D / 4.0) * 2.0 ** -23)
            if form == "fp32":
                vl = f16(v - vh)
                ph = f16(phi)
                pl = f16(phi - ph)
                dropped = ph @ vl + pl @ vh
            else:
                u = phi * 4096.0
                uh = f16(u)
                dropped = e4m3(phi) @ e4m3(4096.0 * (v - vh)) + e4m3(u - uh) @ e4m3(v)
            assert np.all(np.abs(dropped) <= eps), (form, s)
            worst[form] = max(worst[form], float(np.abs(dropped).max() / eps))
    assert worst["fp32"] < 0.5 and worst["fast"] < 0.5       # random signs: far inside the worst-case bound
