#!/usr/bin/env python
"""Benchmark of the RFF maximiser-sampling hot path (BASELINE.json metric) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c4|c4d] [--no-blocks] [--no-cpu]

A step = one fused RFF-sample + per-sample argmax call over the synthetic candidate set (tkbo_rff_argmax; at the tensor-bound
shapes the library runs it as two passes of the same kernel - the hi*hi MMAs over every work item, then the full contraction
over the items that can hold a maximum - with the result bits of a single pass; roofline.two_pass says how many items the
second pass covered).  Default workload c3
(BASELINE.json configs[2], the largest single-GPU configuration): N = 2^24 Hartmann-6 candidates, d = 6, D = 4096
features, S = 256 posterior samples -- STRONG scaling: the 2^24 candidates are split into contiguous 128-row-aligned
shards over the N ranks (one process per GPU, torchrun), the basis is replicated, and each rank's S (value, index) keys
are merged over NVLink peer memory by the kernel's last CTA (no collective call).

Printed JSON (rank 0, one line):
  value         candidate*samples/s of the whole job, inputs resident in HBM (CUDA events around exactly K steps, max over
                ranks); ms_per_step_median = median of the K per-step event intervals
  e2e           the same through the C ABI with HOST buffers (tkbo_rff_argmax_host / _host_peer): every step copies the
                rank's pinned candidates and the float64 basis to the device, runs the one fused launch (which merges the
                ranks' keys itself) and reads the S global results back
  roofline      fused kernel against the tensor pipe: algorithmic flops 2*N*D*S per step, the executed figure (what the
                passes issue in 16-bit-MMA equivalents + the projection), which contraction form ran
  parity        in-run check of the timed kernel against the fp64 oracle on a 2^16-candidate sample, in BOTH contraction
                forms: max |dv| / max|f|, argmax mismatches, tie count (top-2 gap below 1e-4)
  cpu_baseline  the CPU oracle port timed on this host on a bounded sample
  blocks        c2 (configs[1]), c5 = K2 MPES kernel + the device acquisition pipeline, c1 (ascent at the MPES_Forr sizes),
                the fast contraction form at the headline shape, the headline call as a single full-precision pass
                (TKBO_PRUNE=0), and at 8 GPUs c4 (argmax and duel mode)
  run           tiling, L2 note, merge: everything descriptive that is not part of `config` (both arms share `config`)

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
TF_NOTE = "TF 2.1/TFP 0.9/GPflow 2.0.0-rc1 not installable offline"

WORKLOADS = {
    "c2": dict(N=1 << 20, d=2, D=1024, S=64, lo=-1.5, hi=1.5, ls=0.6,
               desc="c2: Six-Hump-Camel 2-D grid 1024x1024, D=1024 RFF, S=64 samples"),
    "c3": dict(N=1 << 24, d=6, D=4096, S=256, lo=0.0, hi=1.0, ls=0.2,
               desc="c3: Hartmann-6 maximiser sampling, N=2^24 candidates, D=4096 RFF, S=256 samples, candidates sharded over the GPUs"),
    "c4": dict(N=1 << 22, d=4, D=2048, S=4096, lo=0.0, hi=1.0, ls=0.2,
               desc="c4: DTS batch Thompson sampling, N=2^22 duel candidates, D=2048 RFF, S=4096 samples, candidates sharded over the GPUs"),
}
L2_BYTES = 126 << 20


def config_of(name):
    w = WORKLOADS[name]
    return {"workload": w["desc"], "N_total": w["N"], "d": w["d"], "D": w["D"], "S": w["S"]}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "sm_max_mhz": p.get("sm_max_mhz", 1965.0), "source": "MEASURED_PEAKS.json (measured)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "sm_max_mhz": 1965.0,
            "source": "B200_PROFILING.md fallback"}


def ncu_facts():
    """Per-launch DRAM traffic / pipe utilisation of the dominant kernels, extracted from the committed ncu captures of
    this same command by tools/ncu_summary.py --json (profiles/r02_ncu_facts.json); absent -> the fields stay null."""
    path = os.path.join(ROOT, "profiles", "r02_ncu_facts.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh)
    return {}


def make_basis_arrays(D, d, S, ls, lo, hi, seed=0, n_train=41, return_q=False):
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
    return (W, b, Theta, Q) if return_q else (W, b, Theta)


def candidates_host(name, lo_row, hi_row, variant=0):
    """Rows [lo_row, hi_row) of the workload's candidate set.  c2: the 1024x1024 uniform grid (i-major like dts.uniform_grid),
    jittered per variant; others: uniform random points, generated in 2^20-row blocks seeded by block index so that every
    rank draws exactly its own rows of the same global set."""
    w = WORKLOADS[name]
    N, d, lo, hi = w["N"], w["d"], w["lo"], w["hi"]
    if name == "c2":
        rng = np.random.default_rng(1000 + variant)
        g = np.linspace(lo, hi, 1024)
        X = np.stack(np.meshgrid(g, g, indexing="xy"), axis=-1).reshape(-1, 2)
        X = X + rng.uniform(-0.5, 0.5, size=(1, 2)) * (hi - lo) / 1023.0
        return np.clip(X, lo, hi).astype(np.float32)[lo_row:hi_row]
    out = np.empty((hi_row - lo_row, d), dtype=np.float32)
    blk = 1 << 20
    for bi in range(lo_row // blk, (hi_row + blk - 1) // blk):
        rng = np.random.default_rng([2000 + variant, bi])
        rows = rng.uniform(lo, hi, size=(min(blk, N - bi * blk), d)).astype(np.float32)
        a, b_ = max(lo_row, bi * blk), min(hi_row, (bi + 1) * blk)
        out[a - lo_row:b_ - lo_row] = rows[a - bi * blk:b_ - bi * blk]
    return out


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

    def window(self, t0, t1):
        rows = [r for t, r in self.rows if t0 - 0.05 <= t <= t1 + 0.05] or [r for _, r in self.rows]
        sm, mx, pw, reasons = [], [], [], set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": statistics.median(pw) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        return self.window(t0 if t0 is not None else 0.0, t1 if t1 is not None else time.time())


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
        v, ix, _ = rff_oracle.rff_argmax_shared(X[lo:hi], W, b, Theta, chunk=2048, idx_offset=lo)
        return v, ix

    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = [p for p in ex.map(work, range(threads)) if p is not None]
    vals = np.stack([p[0] for p in parts])
    idxs = np.stack([p[1] for p in parts])
    best = vals.max(axis=0)
    idx = np.where(vals == best[None], idxs, np.iinfo(np.int64).max).min(axis=0)
    return best, idx


def blas_single_thread():
    import contextlib
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=1)          # one BLAS thread per worker; workers cover the cores
    except Exception:
        return contextlib.nullcontext()


def time_cpu(name, n_sample, steps, warmup):
    w = WORKLOADS[name]
    threads = os.cpu_count() or 1
    W, b, Theta = make_basis_arrays(w["D"], w["d"], w["S"], w["ls"], w["lo"], w["hi"])
    X = candidates_host(name, 0, n_sample).astype(np.float64)
    with blas_single_thread():
        for _ in range(warmup):
            cpu_argmax_threads(X[: max(2048, n_sample // 8)], W, b, Theta, threads)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_argmax_threads(X, W, b, Theta, threads)
        dt = (time.perf_counter() - t0) / steps
    return {"value": n_sample * w["S"] / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_sample} of the {w['N']} candidates per step (fp64 NumPy oracle port of ff.py:20-21,162-167, shared basis, "
                      f"{threads} worker threads); {TF_NOTE}",
            "ms_per_step": dt * 1e3}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload if args.workload in WORKLOADS else "c3"
    w = WORKLOADS[name]
    # bounded sample: calibrate on a few thousand rows, then size each step so that K steps take about --budget-s (60 s)
    probe_rows = 1 << (14 if w["D"] * w["S"] <= (1 << 17) else 11)
    probe = time_cpu(name, probe_rows, 1, 1)
    rows_per_s = probe["value"] / w["S"]
    steps = max(1, args.steps)
    n_sample = int(min(w["N"], max(1 << 10, rows_per_s * args.budget_s / steps)))
    n_sample -= n_sample % 128
    res = time_cpu(name, n_sample, steps, max(0, min(args.warmup, 1)))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(name),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- K1 helpers (GPU arm)
class Ctx:
    """torch / package / rank bookkeeping shared by the blocks."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
        self.nat = self.pkg._native
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
            warm = torch.zeros(64, dtype=torch.int64, device=self.dev)     # communicator set-up (not a step)
            for _ in range(10):
                dist.all_reduce(warm, op=dist.ReduceOp.MAX)
            torch.cuda.synchronize()
        self.h = self.nat.handle(self.local)
        self.sampler = ClockSampler(self.local) if self.rank == 0 else None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.world == 1:
            return float(x)
        t = self.torch.tensor([float(x)], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def launches(self):
        return int(self.nat.lib().tkbo_launch_count(self.h))

    def peer_merge(self, S):
        """PeerMerge on every rank, or None everywhere if any rank cannot map peer memory (then: NCCL all_reduce(MAX))."""
        if self.world == 1:
            return None
        peer = None
        if os.environ.get("TKBO_MERGE", "peer") == "peer":
            try:
                peer = self.pkg.distributed.PeerMerge(S, self.dev)
            except Exception as err:
                print(f"[bench] rank {self.rank}: peer-memory merge unavailable ({err}); using the NCCL merge", file=sys.stderr)
        agree = self.torch.tensor([1 if peer is not None else 0], device=self.dev)
        self.dist.all_reduce(agree, op=self.dist.ReduceOp.MIN)
        return peer if int(agree.item()) == 1 else None


def tensor_equivalents(cfg):
    """(16-bit-MMA equivalents per product, projection K): 3 = fp16 hi*hi + hi*lo + lo*hi; 2 = fp16 hi*hi + one e4m3 MMA of
    twice the K at twice the rate carrying both corrections."""
    proj_k = ((6 * cfg["d_template"] + 3 + 15) // 16) * 16
    return (2 if cfg.get("e4m3_corrections") else 3), proj_k


def k1_roofline(cfg, n_rows, D, S, ms, pk, form, facts_key, long_run, two_pass=None):
    mma_eq, proj_k = tensor_equivalents(cfg)
    Dpad = cfg["n_slabs"] * 64
    alg = 2.0 * n_rows * D * S / (ms * 1e-3) / 1e12
    one_pass = 2.0 * n_rows * Dpad * cfg["S_chunk"] * cfg["n_chunks"]            # flops of one MMA equivalent over everything
    proj = 2.0 * n_rows * Dpad * proj_k * cfg["n_chunks"]
    if two_pass:       # pass 1: hi*hi only (1 equivalent) + projection over everything; pass 2: the full form over the listed share
        share = two_pass[0] / max(two_pass[1], 1)
        exe = ((1.0 + share * mma_eq) * one_pass + (1.0 + share) * proj) / (ms * 1e-3) / 1e12
    else:
        exe = (mma_eq * one_pass + proj) / (ms * 1e-3) / 1e12
    cos_rate = float(n_rows) * Dpad * cfg["n_chunks"] / (ms * 1e-3)
    cos_peak = 16.0 * 148 * pk["sm_max_mhz"] * 1e6
    facts = ncu_facts().get(facts_key + ("_pass1" if two_pass else ""), {}) or ncu_facts().get(facts_key, {})
    # per-launch DRAM traffic from the committed ncu capture of the same kernel instance; the capture's launch may cover a
    # different number of candidates (traffic is linear in N: every candidate row is read once), so it is scaled and both
    # figures are given
    traffic = None
    if facts.get("dram_bytes_per_launch") and facts.get("units_per_launch"):
        traffic = facts["dram_bytes_per_launch"] * n_rows / facts["units_per_launch"]
    sustained = pk.get("bf16_tflops_sustained") or pk["bf16_tflops"]
    peak = sustained if long_run else pk["bf16_tflops"]
    return {"bound": "tensor", "achieved": alg, "peak": peak, "unit": "TFLOP/s", "frac": alg / peak,
            "traffic": traffic, "traffic_captured": {"dram_bytes": facts.get("dram_bytes_per_launch"),
                                                     "candidates_in_captured_launch": facts.get("units_per_launch")},
            "kernel": "rff_fused_kernel", "form": form, "cta_pairs": bool(cfg.get("cta_pairs")),
            "flops_per_launch": 2.0 * n_rows * D * S, "launch_ms": ms,
            "peak_kind": ("sustained" if long_run else "burst") + " bf16 cuBLAS, " + pk["source"],
            "frac_of_burst_peak": alg / pk["bf16_tflops"], "executed": exe, "executed_frac": exe / peak,
            "executed_frac_of_burst_peak": exe / pk["bf16_tflops"],
            "two_pass": ({"items_in_full_precision_pass": two_pass[0], "items": two_pass[1],
                          "what": "pass 1 = hi*hi MMAs only over every item (1 MMA equivalent per product), pass 2 = the full "
                                  "contraction over the items that can hold a maximum; same result bits as a single pass"}
                         if two_pass else None),
            "tensor_pipe_active_pct_ncu": facts.get("tensor_pipe_active_pct"), "ncu_source": facts.get("source"),
            "algorithmic_bytes_per_launch": 4.0 * n_rows * cfg["d_template"],
            "mufu": {"cos_per_s": cos_rate, "peak_cos_per_s": cos_peak, "frac": cos_rate / cos_peak},
            "note": f"achieved = algorithmic 2*N*D*S of ONE argmax call on one GPU ({n_rows} candidates) / its CUDA-event duration "
                    f"(median over the timed steps); executed = what the tensor pipe does in 16-bit-MMA equivalents: {mma_eq} per "
                    f"product ({'fp16 hi*hi + one e4m3 MMA for both corrections' if mma_eq == 2 else 'hi*hi + hi*lo + lo*hi, fp32-accurate'}) "
                    f"+ the bf16x3 projection X W^T + b (K = {proj_k}); the kernel issues fp16 / bf16 / e4m3 MMAs, never TF32, so the "
                    "bf16 dense peak is the denominator (SURVEY 8(d)'s TF32 peak does not apply); mufu = the feature map's cos "
                    "count against 16/clk/SM, the binding unit for S <= 96"}


def time_steps(ctx, step, steps, warmup):
    """W warm-up steps, then exactly K steps between barrier + synchronize, one event per step boundary.
    -> (ms per step from the whole region, max over ranks; median per-step ms on this rank; launches; (t0, t1) wall)."""
    torch = ctx.torch
    for i in range(warmup):
        step(i)
    ctx.barrier()
    l0 = ctx.launches()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    t0 = time.time()
    evs[0].record()
    out = None
    for i in range(steps):
        out = step(warmup + i)
        evs[i + 1].record()
    ctx.barrier()
    t1 = time.time()
    total = evs[0].elapsed_time(evs[steps])
    per = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    return ctx.max_over_ranks(total) / steps, statistics.median(per), ctx.launches() - l0, (t0, t1), out


def time_wall(ctx, step, steps, warmup):
    for i in range(warmup):
        step(i)
    ctx.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    ctx.barrier()
    return ctx.max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)


def k1_parity(ctx, name, W, b, Theta, rows):
    """The oracle against the kernel on `rows` candidates of the workload, in both contraction forms: every value
    (store-F epilogue) and the fused argmax.  A mismatching index only counts where the fp64 top-2 gap exceeds 1e-4 of
    max|f| (north star); `ties` = samples whose fp64 top-2 gap is below that."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import rff_oracle
    torch, ff = ctx.torch, ctx.pkg.fourier_features
    w = WORKLOADS[name]
    X = candidates_host(name, 0, rows)
    S = w["S"]
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    bounds = np.linspace(0, rows, threads + 1).astype(int)
    with blas_single_thread(), ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(lambda i: rff_oracle.rff_values_shared(X[bounds[i]:bounds[i + 1]].astype(np.float64), W, b, Theta,
                                                                   chunk=2048), range(threads)))
    Fref = np.concatenate(parts, axis=0)                                   # (rows, S) float64
    t_cpu = time.perf_counter() - t0
    scale = np.abs(Fref).max(axis=0)
    srt = np.sort(Fref, axis=0)
    gap = (srt[-1] - srt[-2]) / scale
    ref_idx = np.argmax(Fref, axis=0)
    out = {"rows": rows, "samples": S, "tolerance": 1e-4,
           "ties_top2_gap_below_tol": int((gap < 1e-4).sum()),
           "oracle": "fp64 NumPy port of ff.py:20-21,162-167 (oracle/rff_oracle.py), pinned to reference-produced golden vectors",
           "oracle_seconds": t_cpu,
           "theta_norm_over_max_f": float((np.linalg.norm(Theta * np.sqrt(2.0 / w["D"]), axis=0) / scale).max())}
    Xd = torch.from_numpy(X).to(ctx.dev)
    for form in ("fp32", "fast"):
        basis = ff.SharedBasis(W, b, Theta, device=ctx.dev, precision=form)
        F = basis.eval(Xd).cpu().numpy().T.astype(np.float64)
        val, idx = basis.argmax(Xd)
        val, idx = val.cpu().numpy().astype(np.float64), idx.cpu().numpy()
        mism = idx != ref_idx
        out[form] = {"max_abs_dv_over_max_f": float((np.abs(F - Fref) / scale).max()),
                     "argmax_value_max_rel": float((np.abs(val - Fref[ref_idx, np.arange(S)]) / scale).max()),
                     "index_mismatches": int(mism.sum()),
                     "index_mismatches_with_gap_above_tol": int((mism & (gap > 1e-4)).sum()),
                     "ok": bool((np.abs(F - Fref) / scale).max() < 1e-4 and not (mism & (gap > 1e-4)).any())}
        del basis
    return out


def k1_block(ctx, name, args, precision, steps, warmup, with_e2e, long_run, parity_rows=0):
    """Time K1 on workload `name`, candidates sharded over the ranks (strong scaling).  Returns the block dict (rank 0)."""
    torch, pkg = ctx.torch, ctx.pkg
    w = WORKLOADS[name]
    N, d, D, S = w["N"], w["d"], w["D"], w["S"]
    W, b, Theta, Q = make_basis_arrays(D, d, S, w["ls"], w["lo"], w["hi"], return_q=True)
    lo_row, hi_row = pkg.distributed.shard_range(N, ctx.rank, ctx.world)
    n_loc = hi_row - lo_row
    f_scale = np.abs(Q).max(axis=0)
    basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=ctx.dev, precision=precision, f_scale=f_scale)
    cfg = basis.config(max(n_loc, 1))
    form = basis.precision
    host = torch.from_numpy(candidates_host(name, lo_row, hi_row)).pin_memory()
    shard_bytes = max(1, n_loc * d * 4)
    n_rot = int(min(16, max(2, -(-2 * L2_BYTES // shard_bytes))))          # rotating buffers: > 2x L2 between reuses
    dev_bufs = [host.to(ctx.dev)]
    for v in range(1, n_rot):                                             # content differs per buffer
        dev_bufs.append((dev_bufs[0] + (v * 1e-4)).clamp_(w["lo"], w["hi"]).contiguous())
    peer = ctx.peer_merge(S)
    merge = "none (one rank)" if ctx.world == 1 else (
        "keys pushed to every rank's peer memory by the kernel's last CTA, which also merges (one launch per rank, no collective call)"
        if peer is not None else "one int64 all_reduce(MAX) (NCCL)")

    def step(i):
        if peer is not None:
            return peer.argmax(basis, dev_bufs[i % n_rot], lo_row)
        if ctx.world > 1:
            return pkg.distributed.sharded_argmax_fast(basis, dev_bufs[i % n_rot], lo_row)
        return basis.argmax(dev_bufs[i % n_rot], idx_offset=lo_row)

    ms, ms_med, launches, (t0, t1), out = time_steps(ctx, step, steps, warmup)
    clocks = ctx.sampler.window(t0, t1) if ctx.sampler else None
    if peer is not None:
        peer.check()
    two_pass = basis.two_pass_items()                  # of the last timed step on this rank (None: single pass)
    blk = None
    if ctx.rank == 0:
        pk = peaks()
        n_one = ctx.pkg.distributed.shard_range(N, 0, ctx.world)
        n_rows0 = n_one[1] - n_one[0]
        blk = {"metric": METRIC, "value": N * S / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "ms_per_step_median": ms_med,
               "steps": steps, "warmup": warmup, "config": config_of(name), "form": form, "gpu_launches": launches,
               "clocks": clocks,
               "roofline": k1_roofline(cfg, n_rows0, D, S, ms_med, pk, form, f"k1_{name}_{form}", long_run, two_pass),
               "run": {"tiling": cfg, "candidates_per_rank": n_rows0,
                       "l2": f"{n_rot} rotating candidate buffers per rank ({n_rot * shard_bytes >> 20} MiB > 2 x 126 MiB L2)",
                       "parallelism": f"candidate shards x{ctx.world} (strong scaling), merge: {merge}",
                       "precision_requested": precision,
                       "checksum": float(out[0].double().sum().item())}}
    if with_e2e:
        # end to end through the C ABI with HOST buffers: pinned candidates of this rank + the float64 basis -> device,
        # ONE fused launch per rank (streams the chunks in, delivers + merges the keys), S global results -> host
        W_h, b_h, T_h = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (W, b, Theta))
        res_val = torch.empty(S, dtype=torch.float32).pin_memory()
        res_idx = torch.empty(S, dtype=torch.int64).pin_memory()
        peer2 = ctx.peer_merge(S)
        ff = pkg.fourier_features

        def e2e_step(i):
            if ctx.world == 1:
                ff.argmax_host(W_h, b_h, T_h, host, device=ctx.local, out_val=res_val, out_idx=res_idx, precision=form)
            elif peer2 is not None:
                ff.argmax_host(W_h, b_h, T_h, host, device=ctx.local, out_val=res_val, out_idx=res_idx, peer=peer2,
                               idx_offset=lo_row, precision=form)
            else:      # no peer mapping on this box: per-rank host call, then one all_reduce(MAX) of keys
                ff.argmax_host(W_h, b_h, T_h, host, device=ctx.local, out_val=res_val, out_idx=res_idx, precision=form)
                key = pkg.distributed.pack_argmax_key(res_val, res_idx + lo_row).to(ctx.dev, non_blocking=True)
                ctx.dist.all_reduce(key, op=ctx.dist.ReduceOp.MAX)
                key.cpu()
            return float(res_val[0])

        e2e_steps = max(3, min(steps, 20))
        e2e_ms = time_wall(ctx, e2e_step, e2e_steps, min(3, warmup))
        # same contraction form as the device-resident loop: it must reproduce that result of the same buffer bit for bit
        same = None
        if ctx.rank == 0:
            ref_v, ref_i = (peer.argmax(basis, dev_bufs[0], lo_row) if peer is not None else
                            (pkg.distributed.sharded_argmax_fast(basis, dev_bufs[0], lo_row) if ctx.world > 1
                             else basis.argmax(dev_bufs[0], idx_offset=lo_row)))
            same = bool(torch.equal(ref_v.cpu(), res_val) and torch.equal(ref_i.cpu(), res_idx))
        elif ctx.world > 1:      # keep the ranks' epochs in lock step
            (peer.argmax(basis, dev_bufs[0], lo_row) if peer is not None else
             pkg.distributed.sharded_argmax_fast(basis, dev_bufs[0], lo_row))
        if blk is not None:
            blk["e2e"] = {"value": N * S / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "steps": e2e_steps,
                          "h2d_bytes_per_step": N * d * 4 + ctx.world * (D * d + D + D * S) * 8,
                          "d2h_bytes_per_step": ctx.world * (S * 12 + 8),
                          "equals_device_resident_result": same,
                          "what": "tkbo_rff_argmax_host" + ("_peer" if ctx.world > 1 and peer2 is not None else "") +
                                  " (C ABI, host buffers), per rank: pinned candidate shard + fp64 basis -> device, operand prep, "
                                  "ONE fused-argmax launch that consumes the candidate chunks while their copies land"
                                  + (", delivers its keys to the peers and merges them in its last CTA" if ctx.world > 1 else "")
                                  + ", S global results -> host; wall clock, max over ranks"}
    if parity_rows and ctx.rank == 0 and not args.no_cpu:
        blk["parity"] = k1_parity(ctx, name, W, b, Theta, parity_rows)
    del dev_bufs, basis
    torch.cuda.empty_cache()
    return blk


# ----------------------------------------------------------------------------------------------- K2 / pipeline block (c5)
def mpes_block(ctx, args):
    torch, pkg, nat, dev = ctx.torch, ctx.pkg, ctx.nat, ctx.dev
    common = importlib.import_module("top-k-ranking-bayesian-optimization_b200.acquisitions._common")
    Sm, Q, C, k, M, dq = 1024, 100000, 5, 3, 20, 6
    pk = peaks()
    # joint RFF posterior samples at the 5e5 query points and the 20 maximisers: the f-samples K2 consumes are drawn by K1
    # (store-F epilogue) exactly as the acquisition pipeline does, not i.i.d. noise
    rngq = np.random.default_rng(4)
    Wq, bq, Tq, Qq = make_basis_arrays(1024, dq, Sm, 0.2, 0.0, 1.0, seed=4, return_q=True)
    bq_basis = pkg.fourier_features.SharedBasis(Wq, bq, Tq, device=dev, precision="auto", f_scale=np.abs(Qq).max(axis=0))
    chi_h = rngq.uniform(0, 1, size=(Q * C, dq)).astype(np.float32)
    # the maximisers x* are posterior maximiser samples, as in MPES_Forr.py:295-306 (sample_maximizers, then I_batch): the
    # arg max of each of the S sampled functions over 2^16 candidates (K1), the M most frequent distinct winners
    cand_h = rngq.uniform(0, 1, size=(1 << 16, dq)).astype(np.float32)
    _, widx = bq_basis.argmax(torch.from_numpy(cand_h).to(dev))
    uniq, counts = np.unique(widx.cpu().numpy(), return_counts=True)
    top = uniq[np.argsort(-counts, kind="stable")][:M]
    xst_h = cand_h[top]
    if xst_h.shape[0] < M:
        xst_h = np.concatenate([xst_h, rngq.uniform(0, 1, size=(M - xst_h.shape[0], dq)).astype(np.float32)])
    chi, xst = torch.from_numpy(chi_h).to(dev), torch.from_numpy(xst_h).to(dev)
    F_chi = torch.empty((Sm, Q * C), dtype=torch.float32, device=dev)
    bq_basis.eval(chi, out=F_chi)
    F_star = bq_basis.eval(xst)
    fchi = F_chi.view(Sm, Q, C)
    for _ in range(3):
        common.mpes_from_device_samples(fchi, F_star, k, nat.MPES_RANK)
    torch.cuda.synchronize()
    reps = 11
    l0 = ctx.launches()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for r in range(reps):
        mi = common.mpes_from_device_samples(fchi, F_star, k, nat.MPES_RANK)
        evs[r + 1].record()
    torch.cuda.synchronize()
    launches = ctx.launches() - l0
    mms = statistics.median(evs[r].elapsed_time(evs[r + 1]) for r in range(reps))
    bytes_alg = 4.0 * Sm * Q * C + 4.0 * Sm + 8.0 * Q
    P = 60
    ops_per_query = Sm * (C + (1 + C + C * (C - 1) // 2) + P * k + P)      # exp + rcp + mul + add per (sample, query): SURVEY 8(d)
    facts = ncu_facts().get("k2_c5", {})
    out = {"metric": "MPES acq evals/s", "value": Q / (mms * 1e-3), "unit": "acq evals/s", "ms_per_call": mms,
           "gpu_launches_per_call": launches / reps,
           "config": {"workload": "c5: 1e5 query sets x 1024 MC samples, k=3 of 5, M=20 maximisers", "Q": Q, "C": C, "k": k,
                      "S": Sm, "M": M},
           "samples": "joint RFF posterior draws (D=1024, d=6) at the 5e5 query points and the 20 maximisers, written by K1's "
                      "store-F epilogue (2.05 GB > L2)",
           "roofline": {"bound": "hbm", "achieved": bytes_alg / (mms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": bytes_alg / (mms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": facts.get("dram_bytes_per_launch"),
                        "ncu_source": facts.get("source"), "kernel": "mpes_all_perms_kernel<5,3>",
                        "algorithmic_bytes_per_launch": bytes_alg,
                        "fp32_issue": {"algorithmic_ops_per_query": ops_per_query,
                                       "achieved_ops_per_s": ops_per_query * Q / (mms * 1e-3),
                                       "peak_ops_per_s": 148 * 128 * pk["sm_max_mhz"] * 1e6,
                                       "frac": ops_per_query * Q / (mms * 1e-3) / (148 * 128 * pk["sm_max_mhz"] * 1e6),
                                       "issue_active_pct_ncu": facts.get("issue_active_pct"),
                                       "note": "SURVEY 8(d)'s second bound: S*(C exp + 16 rcp + P*k mul + P add) lane-operations "
                                               "per query against 148 SMs x 128 lanes x max SM clock; the kernel executes 162 "
                                               "warp-instructions per (sample, 32 queries), 56 of them two-wide (FMUL2 / FFMA2), and "
                                               "21 MUFU (16 lanes/clk/SM = 168 cycles): MUFU, issue and FMA pipe are within 20 % of "
                                               "each other and bind it, not HBM"},
                        "mufu": {"per_sample_query": 21, "achieved_per_s": 21.0 * Sm * Q / (mms * 1e-3),
                                 "peak_per_s": 148 * 16 * pk["sm_max_mhz"] * 1e6,
                                 "frac": 21.0 * Sm * Q / (mms * 1e-3) / (148 * 16 * pk["sm_max_mhz"] * 1e6)}},
           "checksum": float(mi.sum())}
    if not args.no_cpu:
        # CPU leg: the NumPy restatement of rank_pes.I_batch (pinned to golden vectors of the unmodified rank_pes.py) on the
        # first Qc query sets of the SAME fp32 samples, all host threads; doubles as the parity check of the timed kernel
        from concurrent.futures import ThreadPoolExecutor
        from oracle import mpes_oracle
        Qc = 2000
        f_cpu = fchi[:, :Qc, :].double().cpu().numpy()
        s_cpu = F_star.double().cpu().numpy()
        threads = os.cpu_count() or 1
        cuts = np.linspace(0, Qc, threads + 1).astype(int)
        tc0 = time.perf_counter()
        with blas_single_thread(), ThreadPoolExecutor(max_workers=threads) as ex:
            parts = list(ex.map(lambda i: mpes_oracle.rank_mpes_from_samples(f_cpu[:, cuts[i]:cuts[i + 1]], s_cpu, topk=k),
                                range(threads)))
        tc = time.perf_counter() - tc0
        mi_cpu = np.concatenate(parts)
        mi_gpu = mi[:Qc].cpu().numpy()
        groups = np.bincount(np.argmax(s_cpu, axis=1), minlength=M)
        out["cpu_baseline"] = {"value": Qc / tc, "unit": "acq evals/s", "cores": threads, "kind": "port",
                               "sample": f"{Qc} of the {Q} query sets (NumPy restatement of rank_pes.py:10-108, {threads} threads)"}
        out["parity"] = {"queries": Qc, "mi_min": float(mi_cpu.min()), "mi_median": float(np.median(mi_cpu)),
                         "mi_max": float(mi_cpu.max()),
                         "max_rel_diff": float(np.max(np.abs(mi_gpu - mi_cpu) / np.maximum(np.abs(mi_cpu), 1e-300))),
                         "samples_won_per_maximiser": [int(g) for g in groups],
                         "max_abs_diff": float(np.max(np.abs(mi_gpu - mi_cpu))), "tolerance": "1e-4 relative",
                         "ok": bool(np.allclose(mi_gpu, mi_cpu, rtol=1e-4, atol=1e-9))}

    def pipeline():
        bq_basis.eval(chi, out=F_chi)
        F_s = bq_basis.eval(xst)
        mi_q = common.mpes_from_device_samples(F_chi.view(Sm, Q, C), F_s, k, nat.MPES_RANK)
        return torch.argmax(mi_q)

    for _ in range(2):
        pipeline()
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for r in range(reps):
        best = pipeline()
        evs[r + 1].record()
    torch.cuda.synchronize()
    pms = statistics.median(evs[r].elapsed_time(evs[r + 1]) for r in range(reps))
    out["pipeline"] = {"ms_per_call": pms, "value": Q / (pms * 1e-3), "unit": "acq evals/s",
                       "what": f"RFF joint f-samples at 5e5 query points + 20 maximisers (D=1024, d=6, S=1024; K1 store-F, {bq_basis.precision} "
                               "form, 2.05 GB written) -> K2 -> argmax; everything stays on the device",
                       "best_query": int(best)}
    del F_chi, fchi, chi
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------------------------- c1 block (MPES_Forr sizes)
def c1_block(ctx, args):
    """config c1 (experiments/MPES/MPES_Forr.py:70-79, 295-301): count=20, n_init=50, D=1000, d=1, 3000 RMSprop steps, then
    the final eval / argmax / gather; beside it the oracle's loop on the CPU for a bounded number of steps, and the
    reference data-flow (per-sample basis, fp64 NumPy) of the grid evaluation at a reduced N."""
    torch, ff = ctx.torch, ctx.pkg.fourier_features
    from oracle import rff_oracle
    rng = np.random.default_rng(7)
    count, n_init, D, d, steps = 20, 50, 1000, 1, 3000
    W = rng.standard_normal((count, D, d)) / 0.2
    b = rng.uniform(0, 2 * np.pi, (count, D, 1))
    theta = rng.standard_normal((count, D, 1)) * 0.1
    x0 = rng.uniform(0, 1, (count, n_init, d))
    xw, _ = ff.ascend(x0, W, b, theta, 0.0, 1.0, num_steps=50, device=ctx.dev)         # warm-up of both kernels' first launch
    ff.eval_argmax_gather(xw, W, b, theta, device=ctx.dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x, ran = ff.ascend(x0, W, b, theta, 0.0, 1.0, num_steps=steps, device=ctx.dev)
    fvals, idx, maxi = ff.eval_argmax_gather(x, W, b, theta, device=ctx.dev)
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    out = {"config": {"workload": "c1: MPES_Forr defaults, count=20, n_init=50, D=1000, d=1, num_steps=3000 (ff.py:110-169)"},
           "steps_run": int(ran), "ms_total": t_gpu * 1e3, "us_per_step": t_gpu * 1e6 / max(1, ran),
           "what": "tkbo_rff_ascent_batched (chunks of 64 steps, one block per start point, no grid barrier; the global early-stop "
                   "test runs once per chunk and a chunk that contains the stop is repeated for the exact step count) + "
                   "tkbo_rff_eval_batched; host arrays in, wall clock"}
    if not args.no_cpu:
        cpu_steps = 150
        tc0 = time.perf_counter()
        x_cpu, ran_cpu = rff_oracle.rmsprop_rho0_ascent(x0, W, b, theta, 0.0, 1.0, num_steps=cpu_steps)
        tc = time.perf_counter() - tc0
        x_gpu, _ = ff.ascend(x0, W, b, theta, 0.0, 1.0, num_steps=ran_cpu, device=ctx.dev)
        out["cpu_baseline"] = {"us_per_step": tc * 1e6 / max(1, ran_cpu), "steps": int(ran_cpu), "cores": "NumPy default threads",
                               "kind": "port", "sample": f"{ran_cpu} of the 3000 steps (oracle restatement of ff.py:139-157, "
                                                         "pinned to the reference loop's golden iterates)",
                               "max_abs_dx_vs_gpu": float(np.abs(x_gpu.cpu().numpy() - x_cpu).max())}
        # reference data-flow with a basis PER SAMPLE (ff.py:20-21, 162-167 as written): count independent (N x D) feature
        # matrices and GEMVs, fp64 NumPy, at the c2 feature sizes and a reduced N
        Ng, S2, D2, d2 = 1 << 11, 64, 1024, 2
        Wp = rng.standard_normal((S2, D2, d2)) / 0.6
        bp = rng.uniform(0, 2 * np.pi, (S2, D2, 1))
        tp = rng.standard_normal((S2, D2, 1))
        Xg = np.tile(rng.uniform(-1.5, 1.5, (1, Ng, d2)), (S2, 1, 1))
        tc0 = time.perf_counter()
        rff_oracle.eval_argmax_gather(Xg, Wp, bp, tp)
        tc = time.perf_counter() - tc0
        out["cpu_reference_dataflow_per_sample_basis"] = {
            "value": Ng * S2 / tc, "unit": UNIT, "kind": "port",
            "sample": f"N={Ng} (cost is linear in N), d=2, D=1024, S=64, one basis per sample as ff.py:44-51 draws them; "
                      "fp64 NumPy; " + TF_NOTE}
    return out


# ----------------------------------------------------------------------------------------------- duel mode (c4d)
def duel_block(ctx, steps, warmup):
    """Dueling-Thompson batch sampling at the c4 shape: n = 2048 base points in 2-D, all n^2 = 2^22 ordered duels (d = 4)
    generated inside the kernel, D = 2048, S = 4096, soft-Copeland winners per sample.  The j axis of the duels is split
    over the ranks (strong scaling); the exchange is one all_reduce(SUM) of the [S, n] int64 fixed-point sums."""
    torch, pkg, dist = ctx.torch, ctx.pkg, ctx.dist
    n, dobj, D, S, ls = 2048, 2, 2048, 4096, 0.2
    d = 2 * dobj
    W, b, Theta, Q = make_basis_arrays(D, d, S, ls, 0.0, 1.0, return_q=True)
    basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=ctx.dev, precision="auto", f_scale=np.abs(Q).max(axis=0))
    lo, hi = pkg.distributed.duel_shard(n, ctx.rank, ctx.world)
    N_rank = n * (hi - lo)
    cfg = basis.config(max(1, N_rank))
    rng = np.random.default_rng(3)
    Xb_host = torch.from_numpy(rng.uniform(0.0, 1.0, size=(n, dobj)).astype(np.float32)).pin_memory()
    Xb = Xb_host.to(ctx.dev)
    fix = torch.zeros((S, n), dtype=torch.int64, device=ctx.dev)

    def step(_i):
        fix.zero_()
        basis.duel_copeland_partial(Xb, lo, hi, fix)
        if ctx.world > 1:
            dist.all_reduce(fix, op=dist.ReduceOp.SUM)
        return basis.copeland_finish(fix, scores=False)[1]

    ms, ms_med, launches, (t0, t1), win = time_steps(ctx, step, steps, warmup)
    clocks = ctx.sampler.window(t0, t1) if ctx.sampler else None
    W_h, b_h, T_h = (torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (W, b, Theta))
    res = torch.empty(S, dtype=torch.int64).pin_memory()

    def e2e_step(i):
        basis.set(W_h.to(ctx.dev, non_blocking=True), b_h.to(ctx.dev, non_blocking=True), T_h.to(ctx.dev, non_blocking=True),
                  f_scale=None if basis.precision == "fp32" else np.abs(Q).max(axis=0))
        Xd = Xb_host.to(ctx.dev, non_blocking=True)
        fix.zero_()
        basis.duel_copeland_partial(Xd, lo, hi, fix)
        if ctx.world > 1:
            dist.all_reduce(fix, op=dist.ReduceOp.SUM)
        res.copy_(basis.copeland_finish(fix, scores=False)[1], non_blocking=True)
        torch.cuda.current_stream(ctx.dev).synchronize()

    e2e_ms = time_wall(ctx, e2e_step, max(3, min(steps, 10)), 2)
    if ctx.rank != 0:
        return None
    pk = peaks()
    n_tot = n * n
    blk = {"metric": METRIC, "value": n_tot * S / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "ms_per_step_median": ms_med,
           "steps": steps, "warmup": warmup, "form": basis.precision, "gpu_launches": launches, "clocks": clocks,
           "config": {"workload": "c4d: DTS batch Thompson, n=2048 base points (d_obj=2), all 2^22 duels [x_i, x_j] generated in the "
                                  "kernel, D=2048, S=4096, soft-Copeland winners", "N_total": n_tot, "d": d, "D": D, "S": S},
           "roofline": k1_roofline(cfg, N_rank, D, S, ms_med, pk, basis.precision, f"k1_c4d_{basis.precision}", True),
           "run": {"tiling": cfg, "candidates_per_rank": N_rank,
                   "l2": "no per-step input beyond 16 KB of base points; every step rewrites the 64 MiB fixed-point sums",
                   "parallelism": f"duel shards along j x{ctx.world} (strong scaling), one int64 all_reduce(SUM) of [S, n]",
                   "checksum": int(win.sum().item())},
           "e2e": {"value": n_tot * S / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                   "h2d_bytes_per_step": ctx.world * ((D * d + D + D * S) * 8 + n * dobj * 4), "d2h_bytes_per_step": ctx.world * S * 8,
                   "what": "pinned float64 basis + base points -> device, operand prep, fused duel/Copeland kernel, winners -> host"}}
    return blk


# ----------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args):
    ctx = Ctx()
    if args.gpus != ctx.world and ctx.world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={ctx.world}")
    name = args.workload
    if name == "c4d":
        blk = duel_block(ctx, args.steps, args.warmup)
        if ctx.rank == 0:
            line = headline(blk, ctx, args)
            line["cpu_baseline"] = None
            print(json.dumps(line), flush=True)
        finish(ctx)
        return
    head = k1_block(ctx, name, args, "auto", args.steps, args.warmup, with_e2e=True, long_run=(name != "c2"),
                    parity_rows=(1 << 16) if ctx.world == 1 else (1 << 13))
    line = headline(head, ctx, args) if ctx.rank == 0 else None
    blocks = {}
    if not args.no_blocks:
        sub_steps = max(5, min(args.steps, 10))
        # the headline shape in the fast contraction form (fp16 hi*hi + one e4m3 correction MMA): opt-in, ~1e-5 ||theta'||
        fast = k1_block(ctx, name, args, "fast", sub_steps, 3, with_e2e=False, long_run=(name != "c2"))
        if ctx.rank == 0:
            blocks[f"{name}_fast_form"] = fast
        if name != "c2":
            # the same call as ONE pass over everything in full precision (TKBO_PRUNE=0): what the headline would be without the
            # two-pass argmax, and the launch whose tensor-pipe figure profiles/r02_prof_k1_c3_*_summary.txt shows
            prev = os.environ.get("TKBO_PRUNE")
            os.environ["TKBO_PRUNE"] = "0"
            try:
                single = k1_block(ctx, name, args, "auto", sub_steps, 3, with_e2e=False, long_run=True)
            finally:
                if prev is None:
                    os.environ.pop("TKBO_PRUNE", None)
                else:
                    os.environ["TKBO_PRUNE"] = prev
            if ctx.rank == 0:
                blocks[f"{name}_single_pass"] = single
        if ctx.world == 8:
            for other in ("c4",):
                if other != name:
                    b4 = k1_block(ctx, other, args, "auto", sub_steps, 3, with_e2e=False, long_run=True)
                    if ctx.rank == 0:
                        blocks[other] = b4
            bd = duel_block(ctx, sub_steps, 3)
            if ctx.rank == 0:
                blocks["c4d"] = bd
        if ctx.world == 1:
            if name != "c2":
                # (c2 takes 0.36-0.37 ms when the SM clock sits near 1.65 GHz, as it does right after the power-capped c3
                # blocks, and 0.41-0.42 ms at the full 1.965 GHz: more cycles at the higher clock)
                blocks["c2"] = k1_block(ctx, "c2", args, "auto", 20, 5, with_e2e=True, long_run=False,
                                        parity_rows=0 if args.no_cpu else (1 << 16))
            blocks["c5_mpes"] = mpes_block(ctx, args)
            blocks["c1"] = c1_block(ctx, args)
    if ctx.rank == 0:
        line["blocks"] = blocks
        if ctx.world == 1 and not args.no_cpu:
            w = WORKLOADS[name]
            n_cpu = 1 << (18 if w["D"] * w["S"] <= (1 << 17) else 13)
            res = time_cpu(name, min(w["N"], n_cpu), 1, 1)
            line["cpu_baseline"] = {k2: res[k2] for k2 in ("value", "unit", "cores", "kind", "sample")}
        else:
            line["cpu_baseline"] = None
        if ctx.sampler:
            ctx.sampler.stop()
        print(json.dumps(line), flush=True)
    finish(ctx)


def headline(blk, ctx, args):
    """The contract's top-level keys from a K1 block."""
    form = blk["form"]
    dtype = ("f32 (fp16 hi*hi + e4m3 correction tcgen05.mma, fp32 accumulate)" if form == "fast"
             else "f32 (fp16 hi/lo split x3 tcgen05.mma = fp32 emulation, fp32 accumulate)")
    line = {"metric": METRIC, "value": blk["value"], "unit": UNIT, "n_gpus": ctx.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": blk["ms_per_step"], "ms_per_step_median": blk["ms_per_step_median"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": blk["config"], "e2e": blk.get("e2e"), "gpu_launches": blk["gpu_launches"], "clocks": blk["clocks"],
            "roofline": blk["roofline"], "run": blk["run"]}
    if "parity" in blk:
        line["parity"] = blk["parity"]
    line["gpu_launches_note"] = "kernels launched through libtkbo.so by rank 0 inside the timed region (every rank launches the same)"
    return line


def finish(ctx):
    if ctx.sampler and ctx.sampler.proc and ctx.sampler.proc.poll() is None:
        ctx.sampler.proc.terminate()
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + ["c4d"])
    ap.add_argument("--no-blocks", action="store_true", help="headline only (no c2 / c5 / c1 / fast-form / c4 blocks)")
    ap.add_argument("--no-mpes", action="store_true", help=argparse.SUPPRESS)     # older spelling of --no-blocks
    ap.add_argument("--no-cpu", action="store_true", help="skip every CPU leg (oracle parity + cpu baselines)")
    ap.add_argument("--budget-s", type=float, default=60.0, help="--impl reference: CPU seconds the K timed steps should take")
    args = ap.parse_args()
    args.no_blocks = args.no_blocks or args.no_mpes
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
