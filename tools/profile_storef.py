import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import bench
pkg = importlib.import_module("top-k-ranking-bayesian-optimization_b200")
dev = torch.device("cuda", 0)
Sm, Q, C, dq = 1024, 100000, 5, 6
rng = np.random.default_rng(4)
Wq, bq, Tq = bench.make_basis_arrays(1024, dq, Sm, 0.2, 0.0, 1.0, seed=4)
form = sys.argv[1] if len(sys.argv) > 1 else "fast"
basis = pkg.fourier_features.SharedBasis(Wq, bq, Tq, device=dev, precision=form)
chi = torch.from_numpy(rng.uniform(0, 1, size=(Q * C, dq)).astype(np.float32)).to(dev)
for _ in range(2):
    F = basis.eval(chi)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); F = basis.eval(chi); e1.record(); torch.cuda.synchronize()
print("store-F", form, e0.elapsed_time(e1), "ms", basis.config(Q * C))
