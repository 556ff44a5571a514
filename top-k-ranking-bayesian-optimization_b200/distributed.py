"""Multi-GPU use of the path.  K1: the candidate axis shards across ranks, (W, b, Theta) are replicated, and the only
exchange is the merge of each rank's S (value, index) keys (8 S bytes per rank: pushed over NVLink peer memory by the
kernel itself, ``PeerMerge``; or one NCCL all_reduce(MAX) / all-gather; gloo in the CPU tests of the merge logic).
K2: the query axis shards, the posterior samples are global, and callers exchange either one (best MI, index) pair per
rank or the MI shards.  One process per GPU (torchrun).
"""
from __future__ import annotations

import numpy as np


def shard_range(N, rank, world_size):
    """Contiguous row range [lo, hi) of rank's candidate shard; sizes differ by at most one 128-row tile."""
    tiles = (N + 127) // 128
    per, rem = divmod(tiles, world_size)
    lo_t = rank * per + min(rank, rem)
    hi_t = lo_t + per + (1 if rank < rem else 0)
    return min(N, lo_t * 128), min(N, hi_t * 128)


def merge_pairs(vals, idxs):
    """vals, idxs: [G, S] arrays/tensors of per-rank (max value, global arg max; idx < 0 = empty shard).
    -> (val[S], idx[S]): largest value, lowest index on exact ties — np.argmax semantics over the union."""
    import torch
    vals = torch.as_tensor(vals)
    idxs = torch.as_tensor(idxs)
    G, S = vals.shape
    v = vals.clone()
    v[idxs < 0] = -float("inf")
    best = v.max(dim=0).values
    big = torch.iinfo(torch.int64).max
    cand = torch.where((v == best[None, :]) & (idxs >= 0), idxs, torch.full_like(idxs, big))
    idx = cand.min(dim=0).values
    idx = torch.where(idx == big, torch.full_like(idx, -1), idx)
    return best, idx


def allgather_argmax(val, idx, group=None):
    """All ranks contribute (val[S], idx[S]) and receive the merged global pairs.  Works with any
    torch.distributed backend (NCCL on GPUs, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if world == 1:
        return val, idx
    # one collective: pack (float32 value bits, int64 index) into an int64 [2, S] payload
    payload = torch.stack([val.to(torch.float32).view(torch.int32).to(torch.int64), idx.to(torch.int64)])
    gathered = [torch.empty_like(payload) for _ in range(world)]
    dist.all_gather(gathered, payload, group=group)
    allp = torch.stack(gathered)                                        # [G, 2, S]
    vals = allp[:, 0, :].to(torch.int32).view(torch.float32)
    idxs = allp[:, 1, :]
    return merge_pairs(vals, idxs)


EMPTY_KEY = -(1 << 63)


def pack_argmax_key(val, idx):
    """Torch restatement of the device-side merge key (csrc/ptx.cuh ``argmax_key``): signed int64 whose maximum is
    "largest float32 value, then lowest index"; idx < 0 -> EMPTY_KEY.  Any device; used by the CPU tests of the merge."""
    import torch
    bits = val.to(torch.float32).contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    neg = (bits >> 31) & 1
    orderable = torch.where(neg == 1, (~bits) & 0xFFFFFFFF, bits | 0x80000000)        # larger float <-> larger unsigned
    hi = orderable ^ 0x80000000                                                      # signed order in the high word
    hi = torch.where(hi >= (1 << 31), hi - (1 << 32), hi)
    key = hi * (1 << 32) + (0xFFFFFFFF - idx.to(torch.int64))
    return torch.where(idx < 0, torch.full_like(key, EMPTY_KEY), key)


def unpack_argmax_key(key):
    """Inverse of pack_argmax_key -> (val float32, idx int64; NaN / -1 for EMPTY_KEY)."""
    import torch
    hi = key >> 32                                                                   # arithmetic shift: signed high word
    lo = key & 0xFFFFFFFF
    orderable = (hi & 0xFFFFFFFF) ^ 0x80000000
    bits = torch.where(orderable >= (1 << 31), orderable & 0x7FFFFFFF, (~orderable) & 0xFFFFFFFF)
    bits = torch.where(bits >= (1 << 31), bits - (1 << 32), bits)
    val = bits.to(torch.int32).view(torch.float32)
    idx = 0xFFFFFFFF - lo
    empty = key == EMPTY_KEY
    return torch.where(empty, torch.full_like(val, float("nan")), val), torch.where(empty, torch.full_like(idx, -1), idx)


def allreduce_argmax(basis, key, group=None):
    """Merge per-rank key vectors (SharedBasis.argmax_key) with ONE all_reduce(MAX) over int64 and unpack:
    the whole exchange is 8*S bytes per rank, one NCCL launch + one unpack launch."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(key, op=dist.ReduceOp.MAX, group=group)
    return basis.unpack_key(key)


def sharded_argmax_fast(basis, X_local, idx_offset, group=None):
    """K1 on this rank's shard (key output) + all_reduce(MAX).  Needs global indices < 2^32 - 1."""
    import torch
    if X_local.shape[0] == 0:
        key = torch.full((basis.S,), EMPTY_KEY, dtype=torch.int64, device=basis.device)
    else:
        key = basis.argmax_key(X_local, idx_offset=idx_offset)
    return allreduce_argmax(basis, key, group=group)


class PeerMerge:
    """Merge of the per-rank (value, index) keys fused into the kernel over NVLink peer memory: no collective call per
    iteration.  Every rank owns a small buffer that all ranks of the node can store to (torch symmetric memory, or a plain
    tensor for a single rank); the last CTA of K1 on rank r stores its S keys into slot r of every buffer and releases a
    flag, the merge kernel on each rank acquires the flags of its own buffer, takes the per-sample maximum and unpacks
    (the all-gather of 8 S bytes per rank + max, done by the producing kernel).  One NVSwitch node, <= 8 ranks.
    All ranks must call ``argmax`` the same number of times (the epoch counter advances in lock step)."""

    def __init__(self, S, device, group=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _native as nat
        self.S, self.device, self.epoch = int(S), torch.device(device), 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.world > 8:
            raise nat.TkboError("PeerMerge serves one NVSwitch node (<= 8 ranks); use sharded_argmax_fast across nodes")
        words = int(nat.lib().tkbo_peer_buffer_bytes(self.world, self.S)) // 8
        if self.world > 1:
            import torch.distributed._symmetric_memory as symm
            self.buf = symm.empty(words, dtype=torch.int64, device=self.device)
            self.buf.zero_()
            self._hdl = symm.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
            ptrs = [int(q) for q in self._hdl.buffer_ptrs]
            torch.cuda.synchronize(self.device)
            dist.barrier(group)                      # every buffer is zeroed before anybody stores into it
        else:
            self.buf = torch.zeros(words, dtype=torch.int64, device=self.device)
            ptrs = [self.buf.data_ptr()]
        self._ptrs = (ctypes.c_void_p * self.world)(*ptrs)
        self._nat = nat
        self._dead = False
        self._stream = None

    def check(self):
        """Raise if a merge of this rank gave up waiting for a peer (its outputs then hold index -2 / NaN).  The status word
        is host-mapped: it is current once the stream that ran the merge has been synchronised; ``argmax`` also looks at
        it at the start of every call, so a timeout surfaces at the latest on the following call.  After a timeout the
        epoch bookkeeping of the group is void (a late peer may deliver into a slot that has been reused): the object
        refuses further use and the group has to build a new PeerMerge behind a barrier."""
        if self._dead:
            raise self._nat.TkboError("PeerMerge: a previous merge timed out; create a new PeerMerge on all ranks")
        ep = int(self._nat.lib().tkbo_peer_status(self._nat.handle(self.device.index)))
        if ep != 0:
            self._dead = True
            raise self._nat.TkboError(f"PeerMerge: a peer did not deliver its keys for epoch {ep} within ~10 s "
                                      "(index -2 was returned for that call)")

    def _same_stream(self):
        import torch
        st = torch.cuda.current_stream(self.device).cuda_stream
        if self._stream is None:
            self._stream = st
        elif st != self._stream:
            raise self._nat.TkboError("PeerMerge: all calls of one object must be issued on the same CUDA stream "
                                      "(the two-slot epoch scheme orders a rank's launches by stream order)")

    def argmax(self, basis, X_local, idx_offset, fused=True):
        """(val[S], idx[S]) over all ranks' shards; X_local are rows [idx_offset, idx_offset + N_local); an empty shard
        delivers "no candidate" so that the peers do not wait for it.  ``fused`` (default): ONE launch per call - the
        kernel's last CTA delivers the keys and merges; otherwise delivery + the separate ``tkbo_merge_peer_keys`` launch."""
        import ctypes
        import torch
        nat = self._nat
        assert basis.S == self.S
        self.check()
        self._same_stream()
        self.epoch += 1
        val = torch.empty(self.S, dtype=torch.float32, device=self.device)
        idx = torch.empty(self.S, dtype=torch.int64, device=self.device)
        if X_local.shape[0] == 0:
            nat.check(nat.lib().tkbo_peer_push_keys(nat.handle(self.device.index), None, self.S, self._ptrs, self.world,
                                                    self.rank, self.epoch, nat.current_stream_ptr(self.device)),
                      "tkbo_peer_push_keys")
        elif fused:
            basis.argmax_peer(X_local, idx_offset, self._ptrs, self.rank, self.epoch, val, idx)
            return val, idx
        else:
            basis.argmax_peer(X_local, idx_offset, self._ptrs, self.rank, self.epoch)
        nat.check(nat.lib().tkbo_merge_peer_keys(nat.handle(self.device.index), ctypes.c_void_p(self.buf.data_ptr()), self.world,
                                                 self.S, self.epoch, nat.ptr(val), nat.ptr(idx),
                                                 nat.current_stream_ptr(self.device)), "tkbo_merge_peer_keys")
        return val, idx


def sharded_argmax(basis, X_local, idx_offset, group=None):
    """K1 on this rank's shard, then the all-gather merge.  ``basis`` is a fourier_features.SharedBasis holding
    the same (W, b, Theta) on every rank; X_local (N_local, d) are rows [idx_offset, idx_offset + N_local)."""
    import torch
    if X_local.shape[0] == 0:                       # more ranks than 128-row tiles: this shard is empty
        val = torch.full((basis.S,), -float("inf"), dtype=torch.float32, device=basis.device)
        idx = torch.full((basis.S,), -1, dtype=torch.int64, device=basis.device)
    else:
        val, idx = basis.argmax(X_local, idx_offset=idx_offset)
    return allgather_argmax(val, idx, group=group)


def duel_shard(n, rank, world_size):
    """[j_begin, j_end) of the second duel argument x_j handled by ``rank`` (every rank keeps all n values of i)."""
    per, rem = divmod(n, world_size)
    lo = rank * per + min(rank, rem)
    return lo, lo + per + (1 if rank < rem else 0)


def sharded_duel_copeland(basis, X_base, group=None, scores=True):
    """Dueling-Thompson soft-Copeland with the n^2 duels split over ranks along j: each rank adds its shard's
    fixed-point sigmoid sums (integers, so the reduction order cannot change the result), ONE all_reduce(SUM) over
    int64 [S, n] combines them and every rank finishes to (score[S, n], winner[S])."""
    import torch.distributed as dist
    n = X_base.shape[0]
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        lo, hi = duel_shard(n, dist.get_rank(group), dist.get_world_size(group))
        fix = basis.duel_copeland_partial(X_base, lo, hi)
        dist.all_reduce(fix, op=dist.ReduceOp.SUM, group=group)
    else:
        fix = basis.duel_copeland_partial(X_base, 0, n)
    return basis.copeland_finish(fix, scores=scores)


# ---- K2 (MPES / PES / indifference MI): the query axis shards, the samples are global ------------------------------------
def query_shard(Q, rank, world_size):
    """Contiguous range [lo, hi) of rank's query sets; sizes differ by at most one 32-query tile (K2's work unit)."""
    tiles = (Q + 31) // 32
    per, rem = divmod(tiles, world_size)
    lo_t = rank * per + min(rank, rem)
    hi_t = lo_t + per + (1 if rank < rem else 0)
    return min(Q, lo_t * 32), min(Q, hi_t * 32)


def best_query(mi_local, q_offset, group=None):
    """(max MI, global query index) over all ranks' shards, lowest index on ties: the ``np.argmax(I_vals)`` of
    MPES_Forr.py:309 without gathering the Q values.  One all-gather of a float64 pair per rank; the same result on every
    rank.  ``mi_local`` may be empty (a rank without queries)."""
    import torch
    import torch.distributed as dist
    mi_local = torch.as_tensor(mi_local)
    if mi_local.numel():
        v = mi_local.to(torch.float64).max()
        i = (mi_local.to(torch.float64) == v).nonzero()[0, 0]      # lowest index attaining it, like np.argmax
        pair = torch.stack([v, (i + q_offset).to(torch.float64)])
    else:
        pair = torch.tensor([-float("inf"), -1.0], dtype=torch.float64, device=mi_local.device)
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return float(pair[0]), int(pair[1])
    gathered = [torch.empty_like(pair) for _ in range(dist.get_world_size(group))]
    dist.all_gather(gathered, pair, group=group)
    allp = torch.stack(gathered).cpu()
    best = allp[:, 0].max()
    idx = allp[(allp[:, 0] == best) & (allp[:, 1] >= 0), 1]
    return float(best), (int(idx.min()) if idx.numel() else -1)


def allgather_queries(mi_local, Q, group=None):
    """Full (Q,) MI vector on every rank from the per-rank shards laid out by ``query_shard`` (what a caller that wants
    all of ``I_vals`` uses; 8 Q / G bytes per rank)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return mi_local
    world = dist.get_world_size(group)
    sizes = [hi - lo for lo, hi in (query_shard(Q, r, world) for r in range(world))]
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=mi_local.dtype, device=mi_local.device)
    buf[: mi_local.numel()] = mi_local
    gathered = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(gathered, buf, group=group)
    return torch.cat([g[:n] for g, n in zip(gathered, sizes)])


def sharded_mpes(fchi_local, fstar, topk, mode, q_offset, group=None, perms=None):
    """K2 on this rank's query shard: fchi_local (S, Q_local, C) are the f-samples at the shard's query sets, fstar
    (S, M) the samples at the maximisers -- the SAME S joint posterior draws on every rank (replicated basis /
    ``broadcast_basis_arrays``), so the grouping by arg max and p(x*) are global without any exchange.
    -> (mi_local (Q_local,) float64, (best value, best global query index))."""
    from .acquisitions import _common
    if fchi_local.shape[1] == 0:
        import torch
        mi_local = torch.empty(0, dtype=torch.float64, device=fchi_local.device)
    else:
        mi_local = _common.mpes_from_device_samples(fchi_local, fstar, topk, mode, perms=perms)
    return mi_local, best_query(mi_local, q_offset, group=group)


def broadcast_basis_arrays(W, b, Theta, src=0, group=None, device=None):
    """Replicate the host-drawn basis from rank ``src`` (once per BO iteration; <= 34 MB at the c4 shape)."""
    import torch
    import torch.distributed as dist
    out = []
    for a in (W, b, Theta):
        t = torch.as_tensor(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a
        if device is not None:
            t = t.to(device)
        t = t.contiguous()
        dist.broadcast(t, src=src, group=group)
        out.append(t)
    return out
