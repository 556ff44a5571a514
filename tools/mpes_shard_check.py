#!/usr/bin/env python
"""Multi-GPU check and timing of K2 with the query axis sharded: every rank holds the same S posterior
samples at the maximisers and the f-samples of its own query shard; results against the full single-GPU computation.
    torchrun --nproc-per-node N --master-addr 127.0.0.1 tools/mpes_shard_check.py [Q_per_gpu]"""
import importlib
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
common = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    nat = pkg._native
    # correctness at a size every rank can also compute in full
    S, Q, C, k, M = 512, 20000, 5, 3, 16
    gen = torch.Generator(device=dev).manual_seed(3)                  # same stream on every rank
    fchi = torch.randn((S, Q, C), device=dev, generator=gen)
    fstar = torch.randn((S, M), device=dev, generator=gen)
    full = common.mpes_from_device_samples(fchi, fstar, k, nat.MPES_RANK)
    lo, hi = pkg.distributed.query_shard(Q, rank, world)
    mi, best = pkg.distributed.sharded_mpes(fchi[:, lo:hi, :].contiguous(), fstar, k, nat.MPES_RANK, lo)
    gathered = pkg.distributed.allgather_queries(mi, Q)
    assert torch.equal(gathered, full), f"rank {rank}: gathered MI differs from the single-GPU result"
    assert best == (float(full.max()), int(torch.argmax(full))), f"rank {rank}: best query differs"
    # weak scaling at c5 per GPU: Q_per_gpu query sets x 1024 samples (timed: K2 + the best-query exchange)
    Qg = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    S = 1024
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    fchi = torch.randn((S, Qg, C), device=dev, generator=gen)
    gen2 = torch.Generator(device=dev).manual_seed(7)
    fstar = torch.randn((S, 20), device=dev, generator=gen2)
    for _ in range(3):
        pkg.distributed.sharded_mpes(fchi, fstar, k, nat.MPES_RANK, rank * Qg)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        _, best = pkg.distributed.sharded_mpes(fchi, fstar, k, nat.MPES_RANK, rank * Qg)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"sharded K2 ok on {world} ranks; {Qg} query sets per GPU: {float(t):.3f} ms per call incl. best-query exchange "
              f"= {world * Qg / float(t) * 1e3:.3e} acq evals/s, best = {best}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
