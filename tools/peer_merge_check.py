#!/usr/bin/env python
"""Multi-GPU check of the peer-memory merge (distributed.PeerMerge) against the NCCL all_reduce(MAX) merge and the
single-GPU result:  torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/peer_merge_check.py"""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rng = np.random.default_rng(3)                       # same arrays on every rank
    N, D, d, S = 200000, 512, 4, 96
    X = rng.uniform(0, 1, (N, d)).astype(np.float32)
    W, b, T = rng.standard_normal((D, d)) / 0.3, rng.uniform(0, 6.28, D), rng.standard_normal((D, S))
    basis = pkg.fourier_features.SharedBasis(W, b, T, device=dev)
    pm = pkg.distributed.PeerMerge(S, dev)
    lo, hi = pkg.distributed.shard_range(N, rank, world)
    Xl = torch.from_numpy(X[lo:hi]).to(dev)
    ref_val, ref_idx = basis.argmax(torch.from_numpy(X).to(dev))
    for it in range(20):
        val, idx = pm.argmax(basis, Xl, lo)
        val2, idx2 = pkg.distributed.sharded_argmax_fast(basis, Xl, lo)
        assert torch.equal(val, val2) and torch.equal(idx, idx2), f"rank {rank} iteration {it}: peer merge != NCCL merge"
        assert torch.equal(val, ref_val) and torch.equal(idx, ref_idx), f"rank {rank} iteration {it}: != single-GPU result"
    # the last rank's shard is empty: it delivers "no candidate" and the others must not wait for more
    lo_last, _ = pkg.distributed.shard_range(N, world - 1, world)
    exp_val, exp_idx = basis.argmax(torch.from_numpy(X[:lo_last]).to(dev))
    for it in range(3):
        mine = Xl if rank < world - 1 else Xl[:0]
        val, idx = pm.argmax(basis, mine, lo)
        assert torch.equal(val, exp_val) and torch.equal(idx, exp_idx), f"rank {rank}: empty-shard merge differs"
    # delivery + separate merge launch gives the same bits as the merge in the kernel's last CTA
    for it in range(3):
        val, idx = pm.argmax(basis, Xl, lo, fused=False)
        assert torch.equal(val, ref_val) and torch.equal(idx, ref_idx), f"rank {rank}: two-launch peer merge differs"
    # host-buffer form (tkbo_rff_argmax_host_peer): pinned shard + fp64 basis in, global results out, one launch per rank
    pm2 = pkg.distributed.PeerMerge(S, dev)
    Xh = torch.from_numpy(X[lo:hi].copy()).pin_memory()
    precise = pkg.fourier_features.SharedBasis(W, b, T, device=dev, precision="fp32")      # the host path's form
    pv, pi = precise.argmax(torch.from_numpy(X).to(dev))
    for it in range(4):
        hv, hi_ = pkg.fourier_features.argmax_host(W, b, T, Xh, device=local, peer=pm2, idx_offset=lo)
        assert np.array_equal(hv, pv.cpu().numpy()) and np.array_equal(hi_, pi.cpu().numpy()), f"rank {rank}: host_peer differs"
    torch.cuda.synchronize()
    pm.check()
    pm2.check()
    # timing: merge overhead per iteration, three ways
    out = {}
    for name, fn in (("peer_fused", lambda: pm.argmax(basis, Xl, lo)), ("peer_two_launches", lambda: pm.argmax(basis, Xl, lo, fused=False)),
                     ("nccl", lambda: pkg.distributed.sharded_argmax_fast(basis, Xl, lo))):
        for _ in range(5):
            fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(100):
            fn()
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) / 100
    if rank == 0:
        print(f"peer merge ok on {world} ranks; ms/iteration: {out}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
