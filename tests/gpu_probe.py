#!/usr/bin/env python
"""Bring-up probe for libtkbo.so on a B200 (not a pytest file): raw ctypes calls, NumPy oracle beside them.
Usage: python tests/gpu_probe.py [k1|k2|batched|all]
"""
import ctypes
import importlib
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
nat = importlib.import_module("top-k-ranking-bayesian-optimization_b200._native")
from oracle import mpes_oracle, rff_oracle  # noqa: E402

L = nat.lib()
dev = torch.device("cuda:0")
H = nat.handle(0)
P = nat.ptr


def stream():
    return nat.current_stream_ptr(dev)


def k1_case(N, D, d, S, seed=0, ls=0.5, lo=-1.0, hi=1.0, check_eval=True):
    rng = np.random.default_rng(seed)
    X = rng.uniform(lo, hi, size=(N, d)).astype(np.float32)
    W = rng.standard_normal((D, d)) / ls
    b = rng.uniform(0, 2 * np.pi, D)
    Theta = rng.standard_normal((D, S)) * rng.uniform(0.1, 3.0, size=(1, S))
    basis = ctypes.c_void_p()
    nat.check(L.tkbo_rff_basis_create(H, ctypes.byref(basis), D, d, S), "basis_create")
    cfg = (ctypes.c_int32 * 16)()
    L.tkbo_rff_basis_config(H, basis, N, cfg)
    tW, tb, tT = (torch.from_numpy(a).to(dev) for a in (W, b, Theta))
    tX = torch.from_numpy(X).to(dev)
    nat.check(L.tkbo_rff_basis_set(H, basis, P(tW), P(tb), P(tT), 0, stream()), "basis_set")
    out = {}
    Fref = rff_oracle.rff_values_shared(X.astype(np.float64), W, b, Theta)
    scale = np.abs(Fref).max(axis=0)
    if check_eval:
        F = torch.full((S, N), float("nan"), dtype=torch.float32, device=dev)
        nat.check(L.tkbo_rff_eval(H, basis, P(tX), N, P(F), N, stream()), "rff_eval")
        torch.cuda.synchronize()
        Fh = F.cpu().numpy().T.astype(np.float64)
        err = np.abs(Fh - Fref) / scale
        out["eval_max_rel"] = float(np.nanmax(err))
        out["eval_nan"] = int(np.isnan(Fh).sum())
    val = torch.empty(S, dtype=torch.float32, device=dev)
    idx = torch.empty(S, dtype=torch.int64, device=dev)
    nat.check(L.tkbo_rff_argmax(H, basis, P(tX), N, 1000, P(val), P(idx), stream()), "rff_argmax")
    torch.cuda.synchronize()
    ref_idx = Fref.argmax(axis=0)
    got_idx = idx.cpu().numpy() - 1000
    ref_val = Fref.max(axis=0)
    got_val = val.cpu().numpy().astype(np.float64)
    out["argmax_val_rel"] = float(np.max(np.abs(got_val - ref_val) / scale))
    mism = got_idx != ref_idx
    # a mismatch is a tie if the fp64 value at the returned index is within 1e-4*scale of the max
    ok_idx = (got_idx >= 0) & (got_idx < N)
    gap = np.where(ok_idx, (ref_val - Fref[np.clip(got_idx, 0, N - 1), np.arange(S)]) / scale, np.inf)
    out["idx_mismatch"] = int(mism.sum())
    out["idx_bad"] = int((mism & (gap > 1e-4)).sum())
    out["cfg"] = list(cfg)[:13]
    nat.check(L.tkbo_rff_basis_destroy(H, basis), "basis_destroy")
    return out


def k1_timing(N, D, d, S, iters=5):
    rng = np.random.default_rng(1)
    W = rng.standard_normal((D, d)) / 0.5
    b = rng.uniform(0, 2 * np.pi, D)
    Theta = rng.standard_normal((D, S))
    tX = torch.rand((N, d), dtype=torch.float32, device=dev)
    basis = ctypes.c_void_p()
    nat.check(L.tkbo_rff_basis_create(H, ctypes.byref(basis), D, d, S), "basis_create")
    tW, tb, tT = (torch.from_numpy(a).to(dev) for a in (W, b, Theta))
    nat.check(L.tkbo_rff_basis_set(H, basis, P(tW), P(tb), P(tT), 0, stream()), "basis_set")
    val = torch.empty(S, dtype=torch.float32, device=dev)
    idx = torch.empty(S, dtype=torch.int64, device=dev)
    for _ in range(2):
        nat.check(L.tkbo_rff_argmax(H, basis, P(tX), N, 0, P(val), P(idx), stream()), "rff_argmax")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        nat.check(L.tkbo_rff_argmax(H, basis, P(tX), N, 0, P(val), P(idx), stream()), "rff_argmax")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    L.tkbo_rff_basis_destroy(H, basis)
    return ms, N * S / (ms * 1e-3), 2.0 * N * D * S / (ms * 1e-3) / 1e12


def k2_case(S, Q, C, k, M, seed=0, mode=0):
    rng = np.random.default_rng(seed)
    basis = rng.standard_normal((Q * C + M, 6))
    fs = (rng.standard_normal((S, 6)) @ basis.T + 0.3 * rng.standard_normal((S, Q * C + M))).astype(np.float32)
    fchi = np.ascontiguousarray(fs[:, :Q * C].reshape(S, Q, C))
    fstar = np.ascontiguousarray(fs[:, Q * C:])
    tchi, tstar = torch.from_numpy(fchi).to(dev), torch.from_numpy(fstar).to(dev)
    arg = torch.empty(S, dtype=torch.int32, device=dev)
    nat.check(L.tkbo_argmax_rows(H, P(tstar), S, M, P(arg), stream()), "argmax_rows")
    mi = torch.empty(Q, dtype=torch.float64, device=dev)
    nat.check(L.tkbo_mpes_topk(H, P(tchi), P(arg), S, Q, C, M, None, 0, k, mode, P(mi), stream()), "mpes_topk")
    torch.cuda.synchronize()
    assert np.array_equal(arg.cpu().numpy(), fstar.argmax(axis=1))
    if mode == 0:
        ref = mpes_oracle.rank_mpes_from_samples(fchi.astype(np.float64), fstar.astype(np.float64), topk=k)
    else:
        ref = np.array([mpes_oracle.pes_I_from_samples(np.concatenate([fstar, fchi[:, q, :]], axis=1).astype(np.float64), M)
                        for q in range(Q)])
    got = mi.cpu().numpy()
    out = {"max_rel": float(np.max(np.abs(got - ref) / np.maximum(np.abs(ref), 1e-12))),
           "max_abs": float(np.max(np.abs(got - ref))), "mi_range": (float(ref.min()), float(ref.max()))}
    # explicit-table kernel on the same data
    perms = torch.from_numpy(mpes_oracle.all_permutations_k_in_n(C, k).astype(np.int32)).to(dev)
    if mode == 0:
        mi2 = torch.empty(Q, dtype=torch.float64, device=dev)
        nat.check(L.tkbo_mpes_topk(H, P(tchi), P(arg), S, Q, C, M, P(perms), perms.shape[0], k, 0, P(mi2), stream()), "mpes_table")
        torch.cuda.synchronize()
        out["table_max_rel"] = float(np.max(np.abs(mi2.cpu().numpy() - ref) / np.maximum(np.abs(ref), 1e-12)))
    return out


def k2_timing(S, Q, C, k, M, iters=3):
    tchi = torch.randn((S, Q, C), dtype=torch.float32, device=dev)
    arg = torch.randint(0, M, (S,), dtype=torch.int32, device=dev)
    mi = torch.empty(Q, dtype=torch.float64, device=dev)
    for _ in range(2):
        nat.check(L.tkbo_mpes_topk(H, P(tchi), P(arg), S, Q, C, M, None, 0, k, 0, P(mi), stream()), "mpes_topk")
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        nat.check(L.tkbo_mpes_topk(H, P(tchi), P(arg), S, Q, C, M, None, 0, k, 0, P(mi), stream()), "mpes_topk")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return ms, Q / (ms * 1e-3), 4.0 * S * Q * C / (ms * 1e-3) / 1e9


def batched_case():
    rng = np.random.default_rng(5)
    count, n, D, d = 20, 50, 1000, 1
    X = rng.uniform(0, 1, (count, n, d))
    W = rng.standard_normal((count, D, d)) / 0.2
    b = rng.uniform(0, 2 * np.pi, (count, D))
    th = rng.standard_normal((count, D))
    tX, tW, tb, tt = (torch.from_numpy(a).to(dev) for a in (X, W, b, th))
    f = torch.empty((count, n), dtype=torch.float64, device=dev)
    idx = torch.empty(count, dtype=torch.int64, device=dev)
    nat.check(L.tkbo_rff_eval_batched(H, P(tX), P(tW), P(tb), P(tt), count, n, D, d, P(f), P(idx), stream()), "eval_batched")
    torch.cuda.synchronize()
    fref, iref, _ = rff_oracle.eval_argmax_gather(X, W, b[:, :, None], th[:, :, None])
    out = {"f_max_abs": float(np.abs(f.cpu().numpy() - fref).max()), "idx_equal": bool(np.array_equal(idx.cpu().numpy(), iref))}
    phi = torch.empty((count, n, D), dtype=torch.float64, device=dev)
    nat.check(L.tkbo_rff_features_batched(H, P(tX), P(tW), P(tb), count, n, D, d, P(phi), stream()), "features")
    torch.cuda.synchronize()
    out["phi_max_abs"] = float(np.abs(phi.cpu().numpy() - rff_oracle.fourier_features(X, W, b[:, :, None])).max())
    # ascent
    steps = ctypes.c_int(0)
    xs = tX.clone()
    t0 = time.time()
    nat.check(L.tkbo_rff_ascent_batched(H, P(xs), P(tW), P(tb), P(tt), count, n, D, d, 0.0, 1.0, 300, 1e-3, 1e-7, 1e-7,
                                        ctypes.byref(steps), stream()), "ascent")
    torch.cuda.synchronize()
    out["ascent_ms"] = (time.time() - t0) * 1e3
    xr, sr = rff_oracle.rmsprop_rho0_ascent(X, W, b[:, :, None], th[:, :, None], 0.0, 1.0, num_steps=300)
    out["ascent_steps"] = (steps.value, sr)
    out["ascent_x_max_abs"] = float(np.abs(xs.cpu().numpy() - xr).max())
    return out


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print("device:", torch.cuda.get_device_name(0), "SMs:", L.tkbo_sm_count(H), flush=True)
    if what == "k1time":
        for (N, D, d, S) in [(1 << 20, 1024, 2, 64), (1 << 18, 4096, 6, 256)]:
            ms, cs, tf = k1_timing(N, D, d, S)
            print(f"K1 timing N={N} D={D} d={d} S={S}: {ms:.3f} ms  {cs:.3e} cand*samples/s  {tf:.1f} TFLOP/s(alg)", flush=True)
    if what in ("k1", "all"):
        for (N, D, d, S) in [(128, 64, 2, 32), (256, 128, 2, 64), (1000, 200, 1, 20), (5000, 1000, 6, 96),
                             (4096, 256, 4, 256), (3000, 192, 3, 300), (2048, 128, 12, 512)]:
            t0 = time.time()
            r = k1_case(N, D, d, S)
            print(f"K1 N={N} D={D} d={d} S={S}: {r}  ({time.time() - t0:.1f}s)", flush=True)
        for (N, D, d, S) in [(1 << 20, 1024, 2, 64), (1 << 18, 4096, 6, 256), (1 << 16, 2048, 4, 4096)]:
            ms, cs, tf = k1_timing(N, D, d, S)
            print(f"K1 timing N={N} D={D} d={d} S={S}: {ms:.3f} ms  {cs:.3e} cand*samples/s  {tf:.1f} TFLOP/s(alg)", flush=True)
    if what == "k2time":
        ms, qs, gbs = k2_timing(1024, 100000, 5, 3, 20)
        print(f"K2 timing c5: {ms:.3f} ms  {qs:.3e} evals/s  {gbs:.1f} GB/s", flush=True)
    if what in ("k2", "all"):
        for (S, Q, C, k, M, mode) in [(512, 70, 5, 3, 20, 0), (256, 33, 4, 4, 6, 0), (300, 40, 3, 1, 5, 0),
                                      (200, 10, 5, 2, 40, 0), (1000, 12, 2, 1, 20, 1), (128, 9, 6, 3, 4, 0)]:
            print(f"K2 S={S} Q={Q} C={C} k={k} M={M} mode={mode}: {k2_case(S, Q, C, k, M, mode=mode)}", flush=True)
        ms, qs, gbs = k2_timing(1024, 100000, 5, 3, 20)
        print(f"K2 timing c5: {ms:.3f} ms  {qs:.3e} evals/s  {gbs:.1f} GB/s", flush=True)
    if what in ("batched", "all"):
        print("batched:", batched_case(), flush=True)


if __name__ == "__main__":
    main()
