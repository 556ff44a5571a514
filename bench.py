#!/usr/bin/env python
"""Benchmark of the RFF maximiser-sampling hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3s|c4s|c4d] [--no-mpes]

A step = one fused RFF-sample + per-sample argmax pass over one batch of synthetic candidates (workload c2 by
default: configs[1], N = 2^20 Six-Hump-Camel grid points, d = 2, D = 1024 features, S = 64 posterior samples),
followed for N > 1 GPUs by the all-gather merge of the (value, index) pairs.  Multi-GPU runs shard the candidate
axis: every rank owns its own 2^20-candidate shard (weak scaling), launched by torchrun, one rank per GPU, NCCL.

Printed JSON (rank 0, one line): `value` = candidate*samples/s with inputs resident in HBM (CUDA events, max over
ranks); `e2e` = the same through the public API with HOST buffers (pinned host -> device copy of the step's
candidates and basis, kernel, device -> host read of the result, every step); `roofline` for the fused kernel
(tensor pipe; algorithmic flops 2*N*D*S per launch, plus the executed figure: the fp32-emulating split issues 3
fp16 MMAs per product); `cpu_baseline` = the CPU oracle port timed on this host on a bounded sample;
`mpes` = the K2 kernel at c5 (10^5 query sets x 1024 MC samples, k = 3 of 5) with its HBM roofline.

`--impl reference` times the CPU implementation of the same path (the fp64 NumPy oracle port of the reference's
data-flow: TF 2.1 / GPflow 2.0.0-rc1 cannot be installed offline) on all host threads, same metric and config.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RFF posterior-sample candidate*samples/s (fused argmax)"
UNIT = "candidate*samples/s"

WORKLOADS = {
    # name: (N per GPU, d, D, S, domain lo, hi, lengthscale, description)
    "c2": (1 << 20, 2, 1024, 64, -1.5, 1.5, 0.6, "c2: Six-Hump-Camel 2-D grid 1024x1024, D=1024 RFF, S=64 samples"),
    "c3s": (1 << 21, 6, 4096, 256, 0.0, 1.0, 0.2, "c3 shard: Hartmann-6, 2^21 of the 2^24 candidates, D=4096, S=256"),
    "c4s": (1 << 19, 4, 2048, 4096, 0.0, 1.0, 0.2, "c4 shard: DTS duels, 2^19 of the 2^22 candidates, D=2048, S=4096"),
}
N_ROTATING = 32     # distinct candidate buffers cycled through so every step reads its candidates from HBM, not L2


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "MEASURED_PEAKS.json (measured)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "B200_PROFILING.md fallback"}


def make_basis_arrays(D, d, S, ls, lo, hi, seed=0, n_train=41):
    """W, b and Theta from the ff.py:56-87 regression on a synthetic fitted posterior; host float64."""
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((D, d)) / ls
    b = rng.uniform(0.0, 2.0 * np.pi, D)
    Xtr = rng.uniform(lo, hi, size=(n_train, d))
    K = np.exp(-0.5 * (((Xtr[:, None] - Xtr[None]) / ls) ** 2).sum(-1))
    L = np.linalg.cholesky(K + 1e-6 * np.eye(n_train))
    q_mu, q_sqrt = L @ rng.standard_normal((n_train, 1)), 0.3 * L
    phi = np.sqrt(2.0 / D) * np.cos(Xtr @ W.T + b)                       # (n, D)
    Q = q_sqrt @ rng.standard_normal((n_train, S)) + q_mu                # (n, S)
    Theta = phi.T @ np.linalg.inv(phi @ phi.T) @ Q                       # n < D branch of ff.py:82
    return W, b, Theta


def candidates_host(N, d, lo, hi, variant):
    """c2: the 1024x1024 uniform grid (i-major like dts.uniform_grid), jittered per variant so the rotating buffers
    differ; other shapes: uniform random points."""
    rng = np.random.default_rng(1000 + variant)
    if d == 2 and N == 1 << 20:
        g = np.linspace(lo, hi, 1024)
        X = np.stack(np.meshgrid(g, g, indexing="xy"), axis=-1).reshape(-1, 2)
        X = X + rng.uniform(-0.5, 0.5, size=(1, 2)) * (hi - lo) / 1023.0
        return np.clip(X, lo, hi).astype(np.float32)
    return rng.uniform(lo, hi, size=(N, d)).astype(np.float32)


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU arm
def cpu_argmax_threads(X, W, b, Theta, threads):
    """The oracle's chunked shared-basis evaluation + argmax (ff.py:20-21, 162-167 with n_init -> N), row ranges
    spread over `threads` workers (NumPy releases the GIL in cos and matmul)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import rff_oracle
    N = X.shape[0]
    bounds = np.linspace(0, N, threads + 1).astype(int)

    def work(i):
        lo, hi = bounds[i], bounds[i + 1]
        if hi <= lo:
            return None
        v, ix, _ = rff_oracle.rff_argmax_shared(X[lo:hi], W, b, Theta, chunk=4096, idx_offset=lo)
        return v, ix

    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = [p for p in ex.map(work, range(threads)) if p is not None]
    vals = np.stack([p[0] for p in parts])
    idxs = np.stack([p[1] for p in parts])
    best = vals.max(axis=0)
    idx = np.where(vals == best[None], idxs, np.iinfo(np.int64).max).min(axis=0)
    return best, idx


def time_cpu(workload, n_sample, steps, warmup):
    N, d, D, S, lo, hi, ls, desc = WORKLOADS[workload]
    threads = os.cpu_count() or 1
    import contextlib
    try:
        from threadpoolctl import threadpool_limits
        limiter = threadpool_limits(limits=1)          # one BLAS thread per worker; workers cover the cores
    except Exception:
        limiter = contextlib.nullcontext()
    W, b, Theta = make_basis_arrays(D, d, S, ls, lo, hi)
    X = candidates_host(N, d, lo, hi, 0)[:n_sample].astype(np.float64)
    with limiter:
        for _ in range(warmup):
            cpu_argmax_threads(X[: max(4096, n_sample // 8)], W, b, Theta, threads)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_argmax_threads(X, W, b, Theta, threads)
        dt = (time.perf_counter() - t0) / steps
    return {"value": n_sample * S / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_sample} of the {N} candidates per step (fp64 NumPy oracle port of ff.py:20-21,162-167, shared basis, "
                      f"{threads} worker threads); TF 2.1/TFP 0.9/GPflow 2.0.0-rc1 not installable offline",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, d, D, S, lo, hi, ls, desc = WORKLOADS[args.workload]
    # bounded sample: calibrate on 2^14 rows, then size each step so that K steps take about a minute
    probe = time_cpu(args.workload, 1 << 14, 1, 1)
    rows_per_s = probe["value"] / S
    steps = max(1, args.steps)
    n_sample = int(min(N, max(1 << 12, rows_per_s * 60.0 / steps)))
    n_sample -= n_sample % 128
    res = time_cpu(args.workload, n_sample, steps, max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "N_per_gpu": N, "d": d, "D": D, "S": S},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    nat = pkg._native
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # communicator set-up (not a step): NCCL connects its channels lazily during the first collectives
        warm = torch.zeros(64, dtype=torch.int64, device=dev)
        for _ in range(20):
            dist.all_reduce(warm, op=dist.ReduceOp.MAX)
        torch.cuda.synchronize()
    N, d, D, S, lo, hi, ls, desc = WORKLOADS[args.workload]
    W, b, Theta = make_basis_arrays(D, d, S, ls, lo, hi)
    basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev)
    cfg = basis.config(N)
    n_rot = N_ROTATING if N * d * 4 * N_ROTATING <= (8 << 30) else max(2, (8 << 30) // (N * d * 4))
    host_bufs = [torch.from_numpy(candidates_host(N, d, lo, hi, world * 0 + v + 100 * rank)).pin_memory()
                 for v in range(min(n_rot, 4))]
    dev_bufs = []
    for v in range(n_rot):
        if v < len(host_bufs):
            dev_bufs.append(host_bufs[v].to(dev))
        else:                                   # further variants: device-side shifted copies (content differs)
            dev_bufs.append((dev_bufs[v % len(host_bufs)] + (v * 1e-4)).clamp_(lo, hi).contiguous())
    offset = rank * N
    h = nat.handle(local)

    # N > 1: the per-rank (value, index) keys are merged over NVLink peer memory by the kernel itself (the last CTA pushes
    # the S keys into every rank's buffer; a tiny merge kernel acquires the flags); TKBO_MERGE=nccl takes the
    # all_reduce(MAX)-over-int64 path instead
    peer = None
    merge = "none"
    if world > 1:
        merge = "one int64 all_reduce(MAX) (NCCL)"
        if os.environ.get("TKBO_MERGE", "peer") == "peer":
            try:
                peer = pkg.distributed.PeerMerge(S, dev)
            except Exception as err:                     # no peer mapping on this box: every rank must take the same path
                print(f"[bench] rank {rank}: peer-memory merge unavailable ({err}); using the NCCL merge", file=sys.stderr)
                peer = None
            agree = torch.tensor([1 if peer is not None else 0], device=dev)
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            if int(agree.item()) == 0:
                peer = None
            else:
                merge = "keys pushed to peer memory by the kernel's last CTA + flag-acquiring merge kernel (no collective call)"

    def step(i):
        if peer is not None:
            return peer.argmax(basis, dev_bufs[i % n_rot], offset)
        if world > 1:       # key output + one all_reduce(MAX) over int64 + unpack
            return pkg.distributed.sharded_argmax_fast(basis, dev_bufs[i % n_rot], offset)
        return basis.argmax(dev_bufs[i % n_rot], idx_offset=offset)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = nat.lib().tkbo_launch_count(h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        out = step(args.warmup + i)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = nat.lib().tkbo_launch_count(h) - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * N * S / (ms_per_step * 1e-3)

    # ---- end to end through the C ABI with HOST buffers (tkbo_rff_argmax_host): every step copies the step's pinned
    # candidates and the float64 basis to the device (one launch consuming the candidate chunks as their copies land), runs the fused
    # argmax and reads the S results back.  N > 1: each rank does that for its shard, then one all_reduce(MAX) of keys.
    W_h, b_h, T_h = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (W, b, Theta))
    res_val = torch.empty(S, dtype=torch.float32).pin_memory()
    res_idx = torch.empty(S, dtype=torch.int64).pin_memory()
    ff = pkg.fourier_features

    def e2e_step(i):
        ff.argmax_host(W_h, b_h, T_h, host_bufs[i % len(host_bufs)], device=local, out_val=res_val, out_idx=res_idx)
        if world > 1:       # merge the per-rank host results: one packed int64 key per sample, all_reduce(MAX), back to the host
            key = pkg.distributed.pack_argmax_key(res_val, res_idx + offset).to(dev, non_blocking=True)
            dist.all_reduce(key, op=dist.ReduceOp.MAX)
            return int(key.cpu()[0])
        return float(res_val[0])

    for i in range(min(3, args.warmup)):
        e2e_step(i)
    barrier()
    e2e_steps = max(3, min(args.steps, 20))
    tw0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - tw0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = N * d * 4 + (D * d + D + D * S) * 8
    d2h = S * (4 + 8)

    line = None
    if rank == 0:
        pk = peaks()
        alg_tf = 2.0 * N * D * S / (ms_per_step * 1e-3) / 1e12        # per GPU: one launch processes N candidates
        proj_k = ((6 * cfg["d_template"] + 3 + 15) // 16) * 16
        Dpad = cfg["n_slabs"] * 64
        # tensor-pipe time in 16-bit-MMA equivalents per product: 3 (fp16 hi*hi + hi*lo + lo*hi) or 2 (fp16 hi*hi + one e4m3 MMA
        # of twice the K at twice the rate, carrying both corrections)
        mma_eq = 2 if cfg.get("e4m3_corrections") else 3
        exec_tf = (mma_eq * 2.0 * N * Dpad * cfg["S_chunk"] * cfg["n_chunks"] + 2.0 * N * Dpad * proj_k * cfg["n_chunks"]) \
            / (ms_per_step * 1e-3) / 1e12
        cos_rate = float(N) * Dpad * cfg["n_chunks"] / (ms_per_step * 1e-3)
        cos_peak = 16.0 * 148 * 1965e6
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": ("f32 (fp16 hi*hi + e4m3 correction tcgen05.mma, fp32 accumulate; <= 1e-4 of max|f|)"
                          if cfg.get("e4m3_corrections") else "f32 (fp16 hi/lo split x3 tcgen05.mma, fp32 accumulate)"),
                "data": "synthetic",
                "config": {"workload": desc, "N_per_gpu": N, "d": d, "D": D, "S": S,
                           "l2": f"{n_rot} rotating candidate buffers ({n_rot * N * d * 4 >> 20} MiB > 126 MiB L2)",
                           "tiling": cfg, "parallelism": f"candidate shards x{world}, merge: {merge}"},
                "e2e": {"value": world * N * S / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "what": "tkbo_rff_argmax_host (C ABI, host buffers): pinned candidates + fp64 basis -> device, operand prep, ONE "
                                "fused-argmax launch that consumes the candidate chunks while their copies land, S results -> host"},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": alg_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": alg_tf / pk["bf16_tflops"], "traffic": 8.75e6 if args.workload == "c2" else None,
                             "kernel": "rff_fused_kernel", "flops_per_launch": 2.0 * N * D * S,
                             "executed": exec_tf, "executed_frac": exec_tf / pk["bf16_tflops"],
                             "mufu": {"cos_per_s": cos_rate, "peak_cos_per_s": cos_peak, "frac": cos_rate / cos_peak,
                                      "note": "the feature map needs N*D*n_chunks cos (MUFU, 16/clk/SM at the max SM clock); "
                                              "this unit, not the tensor pipe, binds small-S shapes such as c2"},
                             "note": "achieved = algorithmic 2*N*D*S per launch / CUDA-event time; executed adds what the "
                                     "tensor pipe really does, in 16-bit-MMA equivalents: " + str(mma_eq) + " per product ("
                                     + ("fp16 hi*hi + one e4m3 MMA for both corrections" if mma_eq == 2 else
                                        "hi*hi + hi*lo + lo*hi, fp32-accurate") + ") and "
                                     "the bf16x3 projection X W^T + b (K = " + str(proj_k) + "); peak = measured bf16 cuBLAS burst ("
                                     + pk["source"] + "); traffic = dram bytes of one launch from the committed ncu capture "
                                     "(profiles/r01_prof_k1_c2_v8_summary.txt; algorithmic: 4*N*d = 8.39e6)"}}
    # ---- MPES kernel at c5
    if rank == 0 and not args.no_mpes:
        Sm, Q, C, k, M = 1024, 100000, 5, 3, 20
        gen = torch.Generator(device=dev).manual_seed(4)
        fchi = torch.randn((Sm, Q, C), device=dev, generator=gen)
        fstar = torch.randn((Sm, M), device=dev, generator=gen)
        common = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
        for _ in range(3):
            common.mpes_from_device_samples(fchi, fstar, k, nat.MPES_RANK)
        torch.cuda.synchronize()
        reps = 9
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        evs[0].record()
        for r in range(reps):
            mi = common.mpes_from_device_samples(fchi, fstar, k, nat.MPES_RANK)
            evs[r + 1].record()
        torch.cuda.synchronize()
        mms = statistics.median(evs[r].elapsed_time(evs[r + 1]) for r in range(reps))     # median of 9 back-to-back calls
        bytes_alg = 4.0 * Sm * Q * C + 4.0 * Sm + 8.0 * Q
        pk = peaks()
        line["mpes"] = {"metric": "MPES acq evals/s", "value": Q / (mms * 1e-3), "unit": "acq evals/s", "ms_per_call": mms,
                        "config": {"workload": "c5: 1e5 query sets x 1024 MC samples, k=3 of 5, M=20 maximisers (f-samples 2.05 GB > L2)"},
                        "roofline": {"bound": "hbm", "achieved": bytes_alg / (mms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                     "frac": bytes_alg / (mms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None,
                                     "kernel": "mpes_all_perms_kernel<5,3>"},
                        "checksum": float(mi.sum())}
        if world == 1 and not args.no_cpu:
            # the reference's algorithm for this kernel on one host core: rank_pes.I_batch restated in NumPy (oracle/,
            # pinned to golden vectors of the unmodified rank_pes.py), on a bounded slice of the same samples; doubles as
            # a parity spot check of the values the timed kernel produced
            from oracle import mpes_oracle
            Qc = 512
            f_cpu = fchi[:, :Qc, :].double().cpu().numpy()
            s_cpu = fstar.double().cpu().numpy()
            tc0 = time.perf_counter()
            mi_cpu = mpes_oracle.rank_mpes_from_samples(f_cpu, s_cpu, topk=k)
            tc = time.perf_counter() - tc0
            mi_gpu = mi[:Qc].cpu().numpy()
            line["mpes"]["cpu_baseline"] = {"value": Qc / tc, "unit": "acq evals/s", "cores": 1, "kind": "port",
                                            "sample": f"{Qc} of the {Q} query sets (NumPy restatement of rank_pes.py:10-108)",
                                            "max_rel_diff_vs_gpu": float(np.max(np.abs(mi_gpu - mi_cpu) / np.abs(mi_cpu)))}
        del fchi, fstar
        # the whole acquisition on the device: RFF joint samples at the query points and the maximisers (K1, store-F
        # epilogue, 1024 samples = 4 chunks of 256), winning maximiser per sample, K2, best query
        dq = 6
        rngq = np.random.default_rng(4)
        Wq, bq, Tq = make_basis_arrays(1024, dq, Sm, 0.2, 0.0, 1.0, seed=4)
        bq_basis = pkg.fourier_features.SharedBasis(Wq, bq, Tq, device=dev)
        chi = torch.from_numpy(rngq.uniform(0, 1, size=(Q * C, dq)).astype(np.float32)).to(dev)
        xst = torch.from_numpy(rngq.uniform(0, 1, size=(M, dq)).astype(np.float32)).to(dev)
        F_chi = torch.empty((Sm, Q * C), dtype=torch.float32, device=dev)

        def pipeline():
            bq_basis.eval(chi, out=F_chi)
            F_star = bq_basis.eval(xst)
            mi_q = common.mpes_from_device_samples(F_chi.view(Sm, Q, C), F_star, k, nat.MPES_RANK)
            return torch.argmax(mi_q)

        for _ in range(2):
            pipeline()
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        for _ in range(reps):
            best = pipeline()
        p1.record()
        torch.cuda.synchronize()
        pms = p0.elapsed_time(p1) / reps
        line["mpes"]["pipeline"] = {"ms_per_call": pms, "value": Q / (pms * 1e-3), "unit": "acq evals/s",
                                    "what": "RFF joint f-samples at 5e5 query points + 20 maximisers (D=1024, d=6, S=1024; K1 store-F, "
                                            "2.05 GB written) -> K2 -> argmax; everything stays on the device",
                                    "best_query": int(best)}
        del F_chi
    # ---- CPU baseline beside it (rank 0, N = 1 only)
    if rank == 0 and world == 1 and not args.no_cpu:
        res = time_cpu(args.workload, min(N, 1 << 18), 1, 1)
        line["cpu_baseline"] = {k2: res[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return out


def run_gpu_duel(args):
    """Workload c4d: Dueling-Thompson batch sampling at the c4 shape.  n = 2048 base points in 2-D, all n^2 = 2^22 ordered
    duels (d = 4) generated inside the kernel, D = 2048 features, S = 4096 posterior samples, soft-Copeland winners per
    sample.  Every rank owns a j-shard of n/8 (the c4 box has 8 GPUs): weak scaling, at 8 ranks the job is exactly c4;
    the exchange is one all_reduce(SUM) of the [S, n] int64 fixed-point sums."""
    import torch
    import torch.distributed as dist
    pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
    nat = pkg._native
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, dobj, D, S, ls = 2048, 2, 2048, 4096, 0.2
    d = 2 * dobj
    W, b, Theta = make_basis_arrays(D, d, S, ls, 0.0, 1.0)
    basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev)
    per = n // 8
    lo, hi = (rank % 8) * per, (rank % 8 + 1) * per
    N_rank = n * per
    cfg = basis.config(N_rank)
    rng = np.random.default_rng(3)
    Xb_host = torch.from_numpy(rng.uniform(0.0, 1.0, size=(n, dobj)).astype(np.float32)).pin_memory()
    Xb = Xb_host.to(dev)
    fix = torch.zeros((S, n), dtype=torch.int64, device=dev)
    h = nat.handle(local)

    def step(_i):
        fix.zero_()
        basis.duel_copeland_partial(Xb, lo, hi, fix)
        if world > 1:
            dist.all_reduce(fix, op=dist.ReduceOp.SUM)
        return basis.copeland_finish(fix, scores=False)[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = nat.lib().tkbo_launch_count(h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    for i in range(args.steps):
        win = step(i)
    e1.record()
    barrier()
    t1 = time.time()
    ms = e0.elapsed_time(e1)
    launches = nat.lib().tkbo_launch_count(h) - launches0
    clocks = sampler.stop(t0, t1) if sampler else None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    # end to end: the float64 basis and the base points come from pinned host memory every step, winners go back
    W_h, b_h, T_h = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (W, b, Theta))
    res = torch.empty(S, dtype=torch.int64).pin_memory()

    def e2e_step(i):
        basis.set(W_h.to(dev, non_blocking=True), b_h.to(dev, non_blocking=True), T_h.to(dev, non_blocking=True))
        Xd = Xb_host.to(dev, non_blocking=True)
        fix.zero_()
        basis.duel_copeland_partial(Xd, lo, hi, fix)
        if world > 1:
            dist.all_reduce(fix, op=dist.ReduceOp.SUM)
        res.copy_(basis.copeland_finish(fix, scores=False)[1], non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()

    for i in range(2):
        e2e_step(i)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    tw0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(i)
    barrier()
    e2e_ms = (time.perf_counter() - tw0) * 1e3 / e2e_steps
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    if rank == 0:
        pk = peaks()
        alg_tf = 2.0 * N_rank * D * S / (ms_per_step * 1e-3) / 1e12
        proj_k = ((6 * cfg["d_template"] + 3 + 15) // 16) * 16
        mma_eq = 2 if cfg.get("e4m3_corrections") else 3
        exec_tf = (mma_eq * 2.0 * N_rank * D * cfg["S_chunk"] * cfg["n_chunks"] + 2.0 * N_rank * D * proj_k * cfg["n_chunks"]) \
            / (ms_per_step * 1e-3) / 1e12
        line = {"metric": METRIC, "value": world * N_rank * S / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": ("f32 (fp16 hi*hi + e4m3 correction tcgen05.mma" if mma_eq == 2 else "f32 (fp16 hi/lo split x3 tcgen05.mma")
                         + ", fp32 accumulate; sigmoid sums in 2^-23 fixed point)",
                "data": "synthetic",
                "config": {"workload": "c4d: DTS batch Thompson, n=2048 base points (d_obj=2), duels [x_i, x_j] generated in the "
                                       "kernel, D=2048, S=4096, soft-Copeland winners; per rank the j-shard of n/8 = 2^19 duels "
                                       "(8 ranks = the full 2^22 duels of c4)",
                           "N_per_gpu": N_rank, "d": d, "D": D, "S": S, "tiling": cfg,
                           "l2": "no per-step input beyond 16 KB of base points; every step rewrites the 64 MiB fixed-point sums",
                           "parallelism": f"duel shards along j x{world}, one int64 all_reduce(SUM) of [S, n]"},
                "e2e": {"value": world * N_rank * S / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": (D * d + D + D * S) * 8 + n * dobj * 4, "d2h_bytes_per_step": S * 8,
                        "what": "pinned float64 basis + base points -> device, operand prep, fused duel/Copeland kernel, winners -> host"},
                "gpu_launches": int(launches), "clocks": clocks,
                "roofline": {"bound": "tensor", "achieved": alg_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                             "frac": alg_tf / pk["bf16_tflops"], "traffic": None, "kernel": "rff_fused_kernel<4, 2, *>",
                             "flops_per_launch": 2.0 * N_rank * D * S, "executed": exec_tf,
                             "executed_frac": exec_tf / pk["bf16_tflops"],
                             "note": "algorithmic 2*N*D*S over the rank's duels; executed = " + str(mma_eq) + " 16-bit-MMA equivalents per product + projection; "
                                     "peak = measured bf16 cuBLAS burst (" + pk["source"] + ")"},
                "cpu_baseline": None, "checksum": int(win.sum().item())}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c4d"])
    ap.add_argument("--no-mpes", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        if args.workload == "c4d":
            raise SystemExit("--impl reference is defined on the argmax workloads (c2, c3s, c4s)")
        run_reference(args)
    elif args.workload == "c4d":
        run_gpu_duel(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
