"""Drop-in for the hot functions of the reference's minn.py.

minn_streaming_metric                <- minn.py:59-112
minn_streaming_metric_parameterized  <- minn.py:697-751
_trailing_average / find_minn_peak   <- minn.py:115-205
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, metric_1d, out
from .core import N_FFT


def minn_streaming_metric_parameterized(rx, symbol_len: int):
    as_np = is_numpy_like(rx)
    arr = np.asarray(rx) if as_np else rx
    batched = arr.ndim == 3
    n = arr.shape[-1]
    if max(n - symbol_len + 1, 0) <= 0 and not batched:     # minn.py:80-82, 723-724
        if as_np:
            return np.zeros(0), np.zeros(0, dtype=complex), np.zeros(0)
        z = torch.zeros(0, device=arr.device)
        return z, torch.zeros(0, dtype=torch.complex64, device=arr.device), z.clone()
    r = engine.metric(arr, "minn", int(symbol_len), want_pr=True, path="tile")
    sq = not batched
    return out(r.M, as_np, sq), out(r.P, as_np, sq), out(r.R, as_np, sq)


def minn_streaming_metric(rx):
    return minn_streaming_metric_parameterized(rx, N_FFT)


def _trailing_average(x, win: int):
    """minn.py:115-128 (computed by the find_minn_peak kernel with a wide-open gate)."""
    as_np = is_numpy_like(x)
    xa = np.asarray(x, dtype=float) if as_np else x
    if win <= 1:
        return xa.copy() if as_np else xa.clone()
    if (xa.size if as_np else xa.numel()) == 0:
        return xa.copy() if as_np else xa.clone()
    # the kernel smooths max(x, 0); shift so that negative inputs survive (average is affine)
    t = metric_1d(xa)
    lo = torch.clamp(t.min(), max=0.0)
    _, _, Ms = engine.find_minn_peak(t - lo + 1.0, smooth_win=win, gate_threshold=0.0, want_ms=True)
    return out(Ms + lo - 1.0, as_np)


def find_minn_peak(M, smooth_win: int = 8, gate_threshold: float = 0.5, search_bounds: tuple[int, int] | None = None):
    as_np = is_numpy_like(M)
    n = np.asarray(M).size if as_np else M.numel()
    if n == 0:
        raise ValueError("Minn metric is empty")                              # minn.py:142-143
    peak, span, Ms = engine.find_minn_peak(metric_1d(M), smooth_win, gate_threshold, search_bounds, want_ms=True)
    pk = int(peak[0].item())
    if pk == -2:
        raise ValueError("Minn metric did not produce a positive peak")       # minn.py:152-153
    s, e = (int(v) for v in span[0].tolist())
    if as_np:
        gate = np.zeros(n, dtype=bool)
        gate[s:e] = True
    else:
        gate = torch.zeros(n, dtype=torch.bool, device=M.device)
        gate[s:e] = True
    return pk, gate, out(Ms, as_np)
