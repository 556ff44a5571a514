"""Shared plumbing for the acquisition modules: move f-samples to the GPU and call K2."""
from __future__ import annotations

import numpy as np

from .. import _native as nat
from ..model import to_numpy


def device_of(device=None):
    torch = nat.require_cuda()
    if device is None:
        return torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    return device if device.index is not None else torch.device("cuda", torch.cuda.current_device())


def as_f32_device(a, dev):
    torch = nat.require_cuda()
    if isinstance(a, torch.Tensor):
        return a.to(device=dev, dtype=torch.float32).contiguous()
    arr = a.numpy() if hasattr(a, "numpy") else np.asarray(a)
    return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float32)).to(dev)


def joint_samples(model, x_all, num_samples, dev, rff_features=None):
    """(S, N) float32 CUDA tensor of joint posterior samples at x_all (N, d).

    Uses ``model.predict_f_samples_device`` when the model offers it (RFF sampling through the fused kernel,
    nothing leaves the GPU); otherwise the reference's call ``model.predict_f_samples(x_all, S)`` -> (S, N, 1)
    (pes.py:106, rank_pes.py:28) and one host->device copy."""
    if rff_features is not None or not hasattr(model, "predict_f_samples"):
        if not hasattr(model, "predict_f_samples_device"):
            raise nat.TkboError("model offers neither predict_f_samples nor predict_f_samples_device")
        kw = {} if rff_features is None else {"D": int(rff_features)}
        return model.predict_f_samples_device(x_all, num_samples, device=dev, **kw)
    fs = model.predict_f_samples(to_numpy(x_all), int(num_samples))
    fs = as_f32_device(fs, dev)
    return fs.reshape(int(num_samples), -1)


def mpes_from_device_samples(fchi, fstar, topk, mode, perms=None):
    """fchi (S, Q, C) and fstar (S, M) float32 CUDA tensors -> MI (Q,) float64 CUDA tensor via tkbo_mpes_topk."""
    torch = nat.require_cuda()
    dev = fchi.device
    S, Q, C = fchi.shape
    M = fstar.shape[1]
    h = nat.handle(dev.index)
    st = nat.current_stream_ptr(dev)
    arg = torch.empty(S, dtype=torch.int32, device=dev)
    nat.check(nat.lib().tkbo_argmax_rows(h, nat.ptr(fstar.contiguous()), S, M, nat.ptr(arg), st), "tkbo_argmax_rows")
    mi = torch.empty(Q, dtype=torch.float64, device=dev)
    if perms is None:
        pptr, P = nat.ptr(None), 0
    else:
        perms = torch.as_tensor(np.ascontiguousarray(perms, dtype=np.int32)).to(dev)
        pptr, P = nat.ptr(perms), int(perms.shape[0])
    nat.check(nat.lib().tkbo_mpes_topk(h, nat.ptr(fchi.contiguous()), nat.ptr(arg), S, Q, C, M, pptr, P, int(topk), int(mode),
                                       nat.ptr(mi), st), "tkbo_mpes_topk")
    return mi
