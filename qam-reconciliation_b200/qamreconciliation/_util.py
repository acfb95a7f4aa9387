"""Host <-> device plumbing shared by the extension-class mirrors (torch is plumbing, not product)."""
import os

import numpy as np
import torch

from . import _abi

_NP = {torch.float64: np.float64, torch.float32: np.float32, torch.int64: np.int64,
       torch.uint8: np.uint8, torch.int32: np.int32}


def device():
    return torch.device("cuda", _abi.require_cuda())


def stream():
    return torch.cuda.current_stream().cuda_stream


def to_dev(x, dtype):
    """Contiguous CUDA tensor of `dtype` from numpy / torch / buffer / list input.

    Read-only buffers are accepted (pandas-3 `to_numpy()` hands out read-only arrays and the
    reference's sims pass them straight in: sims/sim_reconciliation.py:60-61)."""
    dev = device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=dtype).contiguous()
    a = np.asarray(x)
    if a.dtype.kind == "S":            # the reference's format-'c' bit arrays
        a = a.view(np.uint8)
    a = np.array(a, dtype=_NP[dtype], copy=True, order="C")
    return torch.from_numpy(a).to(dev)


def to_np(t):
    return t.detach().cpu().numpy()


def env_precision(default="fp64"):
    p = os.environ.get("QAMRECON_PRECISION", default).lower()
    if p in ("fp64", "f64", "64", "double"):
        return _abi.QR_F64
    if p in ("fp32", "f32", "32", "float"):
        return _abi.QR_F32
    raise ValueError(f"QAMRECON_PRECISION={p!r} (expected fp32 or fp64)")


def dtype_code(t):
    if t.dtype == torch.float64:
        return _abi.QR_F64
    if t.dtype == torch.float32:
        return _abi.QR_F32
    raise ValueError(f"LLR arrays must be float32 or float64, got {t.dtype}")
