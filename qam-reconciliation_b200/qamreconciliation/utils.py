"""Small helpers (reference: qamreconciliation/utils.pyx)."""
import torch

from . import _abi
from ._util import dtype_code, stream, to_dev


def dist_cut(x):
    """utils.pyx:18-23"""
    if x < 0:
        return 0
    if x > 1:
        return 1
    return x


def count_errors_batch(lappr, word, k=None):
    """Per frame: #{i < k : hard decision of lappr[b, i] != word[b, i]} (utils.pyx:27-40), int32[B] on the GPU."""
    if isinstance(lappr, torch.Tensor) and lappr.dtype == torch.float32:
        l = to_dev(lappr, torch.float32)
    else:
        l = to_dev(lappr, torch.float64)
    w = to_dev(word, torch.uint8)
    if l.dim() != 2 or w.shape != l.shape:
        raise ValueError("Sizes do not match")
    if k is None:
        k = l.shape[1]
    err = torch.empty(l.shape[0], dtype=torch.int32, device=l.device)
    _abi.check(_abi.lib().qr_count_errors(l.data_ptr(), dtype_code(l), w.data_ptr(), l.shape[0], l.shape[1], int(k),
                                          err.data_ptr(), stream()))
    return err


def count_errors_from_lappr(lappr, word):
    """utils.pyx:27-40: lappr >= 0 decides bit 0."""
    l = to_dev(lappr, torch.float64).reshape(1, -1)
    w = to_dev(word, torch.uint8).reshape(1, -1)
    if l.shape != w.shape:
        raise ValueError("Sizes do not match")
    return int(count_errors_batch(l, w)[0].item())
