"""Mutual-information Monte Carlo on the GPU (reference: qamreconciliation/mutual_information.pyx:29-39,
:212-300).  The quadrature-based functions of the reference (scipy.integrate.quad,
:43-209) are analysis tools outside the reconciliation path and are not provided."""
import numpy as np
import torch

from . import _abi
from ._util import stream, to_dev
from .noisemapper import _demap_mode


def P_xhat(nm):
    """mutual_information.pyx:29-39: P(xhat = a_i) = sum_j p_j P(xhat = a_i | x = a_j)"""
    p = np.asarray(nm.probabilities, dtype=np.float64)
    f = np.asarray(nm.fwrd_transition_probability, dtype=np.float64)
    res = np.zeros(nm.order)
    for i in range(nm.order):
        for j in range(nm.order):
            res[i] += p[j] * f[j, i]
    return res


def information_from_samples(nm, p_Xhat, x_ind, y, which=(1, 1, 1), mode=None):
    """The three estimates of montecarlo_information (:241-298) for given samples (any shape, flattened):
    x_ind = Alice's symbol indices, y = Bob's channel outputs.  One kernel launch; returns floats."""
    xi = to_dev(x_ind, torch.int64).reshape(-1); yy = to_dev(y, torch.float64).reshape(-1)
    if xi.numel() != yy.numel():
        raise ValueError("Sizes do not match")
    mask = sum(1 << k for k in range(3) if which[k])
    if mask & 4:
        nm._build_grid()
    px = to_dev(np.ascontiguousarray(p_Xhat, dtype=np.float64), torch.float64)
    sums = torch.zeros(3, dtype=torch.float64, device=yy.device)
    _abi.check(_abi.lib().qr_information_sums(nm._h, px.data_ptr(), xi.data_ptr(), yy.data_ptr(), yy.numel(), mask,
                                              _demap_mode(mode) & 1, sums.data_ptr(), stream()))
    out = (sums / max(1, yy.numel())).cpu().numpy()
    return float(out[0]), float(out[1]), float(out[2])


def montecarlo_information(pa, nm, p_Xhat, N, which=np.ones(3, dtype=np.uint8), generator=None, mode=None):
    """montecarlo_information(pa, nm, p_Xhat, N, which) -> (I_X_Xhat, I_X_Y, I_XN_Xhat)  (:212-300).
    Samples are drawn on the device (the reference uses numpy's global RNG, :236-239): symbols with the
    alphabet's probabilities, y = a[x] + sigma * N(0, 1)."""
    dev = torch.device("cuda", _abi.require_cuda())
    probs = torch.as_tensor(np.asarray(pa.probabilities, dtype=np.float64), device=dev)
    x = torch.multinomial(probs, int(N), replacement=True, generator=generator)
    const = torch.as_tensor(np.asarray(pa.constellation, dtype=np.float64), device=dev)
    y = const[x] + nm.noise_sigma * torch.randn(int(N), dtype=torch.float64, device=dev, generator=generator)
    return information_from_samples(nm, p_Xhat, x, y, which, mode)
