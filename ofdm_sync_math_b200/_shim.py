"""Helpers shared by the drop-in modules (sc, minn, park, combined_sc_min, zc, zc_v2, zc_freq, sync_aa, minn_rtl)."""
from __future__ import annotations

import numpy as np
import torch


def is_numpy_like(x) -> bool:
    return not isinstance(x, torch.Tensor)


def out(t: torch.Tensor | None, as_numpy: bool, squeeze: bool = True):
    """Device tensor -> what the reference would have returned (numpy for numpy callers)."""
    if t is None:
        return None
    if squeeze and t.dim() >= 2 and t.shape[0] == 1:
        t = t[0]
    if as_numpy:
        a = t.detach().cpu().numpy()
        if a.dtype == np.uint8:
            a = a.astype(bool)
        return a
    return t


def metric_1d(M) -> torch.Tensor:
    """1-D metric (numpy or torch) -> float64/float32 row tensor [1, n] on the device."""
    if isinstance(M, torch.Tensor):
        t = M
    else:
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(M, dtype=np.float64)))
    if t.dim() != 1:
        raise ValueError("metric must be 1-D")
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    return t.cuda()[None] if not t.is_cuda else t[None]
