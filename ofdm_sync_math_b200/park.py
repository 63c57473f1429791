"""Drop-in for park.park_streaming_metric (park.py:64-114)."""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, out
from .core import N_FFT


def park_streaming_metric(rx):
    as_np = is_numpy_like(rx)
    arr = np.asarray(rx) if as_np else rx
    batched = arr.ndim == 3
    L = arr.shape[-1]
    half = N_FFT // 2
    if (half == 0 or L < 2 * half + 1) and not batched:              # park.py:79-95
        if as_np:
            return np.zeros(0, dtype=int), np.zeros(0, dtype=float), np.zeros(0, dtype=np.complex128), np.zeros(0, dtype=float)
        z = torch.zeros(0, device=arr.device)
        return torch.zeros(0, dtype=torch.int64, device=arr.device), z, torch.zeros(0, dtype=torch.complex64, device=arr.device), z.clone()
    M, P, E = engine.park_metric(arr, N_FFT)
    n = M.shape[-1]
    ds = np.arange(half, half + n) if as_np else torch.arange(half, half + n, device=M.device)
    sq = not batched
    return ds, out(M, as_np, sq), out(P, as_np, sq), out(E, as_np, sq)
