"""Drop-in for the hot functions of the reference's sc.py (same names, arguments, return shapes).

sc_streaming_metric           <- sc.py:42-78
find_plateau_end_from_metric  <- sc.py:81-146
numpy in -> numpy float64/complex128 out (float64 prefix-sum kernel); CUDA torch tensors in ->
torch tensors out.  Extension: a leading frames axis (3-D input) is accepted and kept.
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, metric_1d, out
from .core import N_FFT


def sc_streaming_metric(rx):
    as_np = is_numpy_like(rx)
    arr = np.asarray(rx) if as_np else rx
    batched = arr.ndim == 3
    n = arr.shape[-1]
    if max(n - N_FFT + 1, 0) <= 0 and not batched:          # sc.py:50-52
        if as_np:
            return np.zeros(0), np.zeros(0, dtype=complex), np.zeros(0)
        z = torch.zeros(0, device=arr.device)
        return z, torch.zeros(0, dtype=torch.complex64, device=arr.device), z.clone()
    r = engine.metric(arr, "sc", N_FFT, want_pr=True, path="tile")
    sq = not batched
    return out(r.M, as_np, sq), out(r.P, as_np, sq), out(r.R, as_np, sq)


def find_plateau_end_from_metric(M, cp_len: int, lookahead: int | None = None, smooth_win: int = 8) -> int:
    if (M.size if isinstance(M, np.ndarray) else M.numel() if isinstance(M, torch.Tensor) else len(M)) == 0:
        return 0                                             # sc.py:94-95
    return int(engine.find_plateau_end(metric_1d(M), cp_len, lookahead, smooth_win)[0].item())
