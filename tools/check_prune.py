"""Two-pass (pruned) argmax against the single pass on the bench shapes: the S (value, index) results must be the same bits.
    python tools/check_prune.py [c3|c2|c4 ...]"""
import importlib
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
dev = torch.device("cuda", 0)


def timed(basis, bufs, reps=12):
    for i in range(3):
        basis.argmax(bufs[i % len(bufs)])
    torch.cuda.synchronize()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    evs[0].record()
    for i in range(reps):
        basis.argmax(bufs[i % len(bufs)])
        evs[i + 1].record()
    torch.cuda.synchronize()
    return statistics.median(evs[i].elapsed_time(evs[i + 1]) for i in range(reps))


for name in sys.argv[1:] or ["c3"]:
    if name.startswith("custom:"):
        cn, cd, cD, cS = (int(v) for v in name[7:].split(","))
        bench.WORKLOADS[name] = dict(N=cn, d=cd, D=cD, S=cS, lo=0.0, hi=1.0, ls=0.2, desc=name)
    w = bench.WORKLOADS[name]
    rows = min(w["N"], 1 << 21)
    W, b, Theta = bench.make_basis_arrays(w["D"], w["d"], w["S"], w["ls"], w["lo"], w["hi"])
    X0 = torch.from_numpy(bench.candidates_host(name, 0, rows)).to(dev)
    bufs = [X0] + [(X0 + v * 1e-4).clamp_(w["lo"], w["hi"]).contiguous() for v in range(1, 4)]
    for form in ("fast", "fp32"):
        res = {}
        for prune in ("0", "2"):
            os.environ["TKBO_PRUNE"] = prune
            basis = pkg.fourier_features.SharedBasis(W, b, Theta, device=dev, precision=form)
            ms = timed(basis, bufs)
            out = [tuple(t.clone() for t in basis.argmax(x)) for x in bufs]
            res[prune] = (ms, out)
            del basis
        same = all(torch.equal(a[0], c[0]) and torch.equal(a[1], c[1]) for a, c in zip(res["0"][1], res["2"][1]))
        print(f"{name} {form} rows={rows}: single pass {res['0'][0]:.3f} ms, two passes {res['2'][0]:.3f} ms, bitwise equal: {same}",
              flush=True)
