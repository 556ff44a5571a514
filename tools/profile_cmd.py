"""The launches profiled for profiles/r02_*: one warm-up + one measured launch of each dominant kernel at the bench shapes,
inputs resident on the device (no host path: ncu serialises kernels against copies).
    python tools/profile_cmd.py                      (plain run; prints CUDA-event times)
    ncu --set full ... -k regex:'rff_fused_kernel|mpes_all_perms' python tools/profile_cmd.py
Kernel instances: rff_fused_kernel<6,0,2,false> = c3 shard (2^21 of the 2^24 candidates) fp32 form, <6,0,2,true> = fast form,
<2,0,4,false> = c2, <4,2,2,*> = c4 duel shard; mpes_all_perms_kernel<5,3,16> = c5 (16-byte staged copies)."""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
dev = torch.device("cuda", 0)


def timed(fn, label):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{label}: {e0.elapsed_time(e1):.3f} ms", flush=True)


which = sys.argv[1:] or ["c3", "c2", "c4d", "c5"]
if "c3" in which:
    w = bench.WORKLOADS["c3"]
    W, b, Theta = bench.make_basis_arrays(w["D"], w["d"], w["S"], w["ls"], w["lo"], w["hi"])
    X = torch.from_numpy(bench.candidates_host("c3", 0, 1 << 21)).to(dev)
    for form in ("fp32", "fast"):
        basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev, precision=form)
        timed(lambda: basis.argmax(X), f"K1 c3 shard 2^21 ({form})")
        del basis
if "c2" in which:
    w = bench.WORKLOADS["c2"]
    W, b, Theta = bench.make_basis_arrays(w["D"], w["d"], w["S"], w["ls"], w["lo"], w["hi"])
    X = torch.from_numpy(bench.candidates_host("c2", 0, 1 << 20)).to(dev)
    basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev, precision="fp32")
    timed(lambda: basis.argmax(X), "K1 c2")
    del basis
if "c4d" in which:
    n, dobj, D, S = 2048, 2, 2048, 4096
    W, b, Theta = bench.make_basis_arrays(D, 2 * dobj, S, 0.2, 0.0, 1.0)
    basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev, precision="fast")
    Xb = torch.rand((n, dobj), device=dev)
    fix = torch.zeros((S, n), dtype=torch.int64, device=dev)
    timed(lambda: basis.duel_copeland_partial(Xb, 0, n // 8, fix), "K1 c4 duel shard (fast)")
    del basis, fix
if "c5" in which:
    common = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
    Sm, Q, C, k, M, dq = 1024, 100000, 5, 3, 20, 6
    rng = np.random.default_rng(4)
    Wq, bq, Tq = bench.make_basis_arrays(1024, dq, Sm, 0.2, 0.0, 1.0, seed=4)
    basis = pkg.fourier_features.SharedBasis(Wq, bq, Tq, device=dev, precision="fp32")
    chi = torch.from_numpy(rng.uniform(0, 1, size=(Q * C, dq)).astype(np.float32)).to(dev)
    cand = torch.from_numpy(rng.uniform(0, 1, size=(1 << 16, dq)).astype(np.float32)).to(dev)
    _, widx = basis.argmax(cand)
    uniq, counts = np.unique(widx.cpu().numpy(), return_counts=True)
    xst = cand[torch.from_numpy(uniq[np.argsort(-counts, kind="stable")][:M]).to(dev)]
    F_chi = basis.eval(chi).view(Sm, Q, C)
    F_star = basis.eval(xst)
    timed(lambda: common.mpes_from_device_samples(F_chi, F_star, k, pkg._native.MPES_RANK), "K2 c5")
