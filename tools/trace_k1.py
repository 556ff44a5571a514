#!/usr/bin/env python
"""Pipeline timeline of the fused RFF kernel: CTA 0 logs clock64() at every mbarrier hand-off (tkbo_debug_set_trace),
this script prints per-role per-slab intervals so that bubbles are visible without a profiler.

    python tools/trace_k1.py [N D d S] [first_slab n_slabs]
"""
import ctypes
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("TKBO_LIB", os.path.join(ROOT, "top-k-ranking-bayesian-optimization_b200", "libtkbo_trace.so"))
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
nat = pkg._native

EV = ["tma_empty", "tma_issued", "p0_pfull", "p0_arrived", "p1_pfull", "p1_arrived", "p2_pfull", "p2_arrived",
      "p3_pfull", "p3_arrived", "mma_proj", "mma_afull", "mma_committed", "epi_accfull", "epi_done", "mma_accempty",
      "mma_proj_issued"]
E_PROJ, E_AFULL, E_COMMIT, E_EPIFULL, E_EPIDONE, E_ACCEMPTY, E_PROJISS = 10, 11, 12, 13, 14, 15, 16


def main():
    a = [int(v) for v in sys.argv[1:]]
    N, D, d, S = (a + [1 << 20, 1024, 2, 64])[:4] if len(a) < 4 else a[:4]
    first, count = (a[4], a[5]) if len(a) >= 6 else (160, 24)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(0)
    W = rng.standard_normal((D, d)) / 0.5
    b = rng.uniform(0, 2 * np.pi, D)
    Theta = rng.standard_normal((D, S))
    X = torch.from_numpy(rng.uniform(-1, 1, (N, d)).astype(np.float32)).to(dev)
    basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev)
    cfg = basis.config(N)
    print("config:", cfg)
    for _ in range(3):
        basis.argmax(X)
    torch.cuda.synchronize()
    cap = 4096
    trace = torch.zeros((len(EV), cap), dtype=torch.int64, device=dev)
    h = nat.handle(0)
    nat.check(nat.lib().tkbo_debug_set_trace(h, ctypes.c_void_p(trace.data_ptr()), cap), "set_trace")
    basis.argmax(X)
    torch.cuda.synchronize()
    nat.check(nat.lib().tkbo_debug_set_trace(h, None, 0), "set_trace")
    t = trace.cpu().numpy().astype(np.int64)
    t0 = t[t > 0].min()
    npg = cfg["producer_groups"]
    coop = bool(cfg.get("producers_share_slab"))
    grp = (lambda it: (0, it)) if coop else (lambda it: (it % npg, it // npg))    # producer group 0's log stands for the slab
    n_it = int(np.nonzero(t[E_COMMIT] > 0)[0].max()) + 1
    n_items = int((t[E_EPIFULL] > 0).sum())
    total = t.max() - t0
    print(f"CTA 0: {n_it} slab iterations, {n_items} items, {total} cycles total, {total / max(n_it, 1):.0f} cycles/slab")
    rel = lambda v: int(v - t0) if v > 0 else -1
    print(f"{'it':>5} {'tma_empty':>10} {'tma_iss':>8} {'mma_proj':>9} {'prod_pfull':>10} {'prod_arr':>9} {'prod_dur':>8} "
          f"{'mma_afull':>10} {'mma_commit':>10} {'d_commit':>8}")
    for it in range(first, min(first + count, n_it)):
        g, k = grp(it)
        pb, pa = t[2 + 2 * g][k], t[3 + 2 * g][k]
        dc = t[E_COMMIT][it] - t[E_COMMIT][it - 1] if it > 0 else 0
        print(f"{it:>5} {rel(t[0][it]):>10} {rel(t[1][it]):>8} {rel(t[E_PROJ][it]):>9}/{rel(t[E_PROJISS][it]) - rel(t[E_PROJ][it]):<4} {rel(pb):>10} {rel(pa):>9} {int(pa - pb):>8} "
              f"{rel(t[E_AFULL][it]):>10} {rel(t[E_COMMIT][it]):>10} {int(dc):>8}")
    print("items: epi_accfull, epi_done (dur), mma_accempty")
    for i in range(min(n_items, 6)):
        print(f"{i:>5} {rel(t[E_EPIFULL][i]):>10} {rel(t[E_EPIDONE][i]):>10} ({int(t[E_EPIDONE][i] - t[E_EPIFULL][i])}) {rel(t[E_ACCEMPTY][i]):>10}")
    ii = np.arange(1, n_it)
    ii = ii[(t[E_AFULL][ii] > 0) & (t[E_COMMIT][ii - 1] > 0) & (t[E_PROJISS][ii] > 0)]   # paired bursts log their first slab only
    print(f"proj warp per slab: waits {(t[E_PROJ][ii] - t[E_PROJISS][ii - 1]).mean():.0f} | issue {(t[E_PROJISS][ii] - t[E_PROJ][ii]).mean():.0f};  "
          f"main warp per slab: waits {(t[E_AFULL][ii] - t[E_COMMIT][ii - 1]).mean():.0f} | issue {(t[E_COMMIT][ii] - t[E_AFULL][ii]).mean():.0f}")
    print(f"proj runs ahead of main by {(t[E_AFULL][ii] - t[E_PROJISS][ii]).mean():.0f} cycles")
    wait_a = (t[E_AFULL][ii] - t[E_COMMIT][ii - 1]).clip(min=0)
    issue = t[E_COMMIT][ii] - t[E_AFULL][ii]
    print(f"MMA warp per slab: from previous commit to a_full seen (projection issue + waiting) {wait_a.mean():.0f}, "
          f"main issue+commit {issue.mean():.0f}")
    for g in range(npg):
        k = int((t[3 + 2 * g] > 0).sum())
        dur = t[3 + 2 * g][:k] - t[2 + 2 * g][:k]
        gap = t[2 + 2 * g][1:k] - t[3 + 2 * g][:k - 1]
        print(f"producer group {g}: {k} slabs, busy {dur.mean():.0f} cycles/slab, waiting for p_full {gap.mean():.0f}")
    k = n_it
    print(f"TMA warp per slab: wait b_empty {(t[0][1:k] - t[1][:k - 1]).mean():.0f}, issue {(t[1][:k] - t[0][:k]).mean():.0f}")
    lat = []
    for it in range(k):
        g, kk = grp(it)
        lat.append(t[2 + 2 * g][kk] - t[E_PROJISS][it])
    print(f"projection issued -> producer sees p_full: mean {np.mean(lat):.0f}, min {np.min(lat)}, max {np.max(lat)}")
    lat2 = [t[E_AFULL][it] - t[3 + 2 * grp(it)[0]][grp(it)[1]] for it in range(k) if t[E_AFULL][it] > 0]
    print(f"producer arrival -> MMA warp sees a_full: mean {np.mean(lat2):.0f}, min {np.min(lat2)}, max {np.max(lat2)}")


if __name__ == "__main__":
    main()
