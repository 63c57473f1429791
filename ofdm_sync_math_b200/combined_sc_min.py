"""Drop-in for the hot functions of the reference's combined_sc_min.py.

minn_streaming_metric          <- combined_sc_min.py:60-113 (identical to minn.py's)
schmidl_cox_streaming_metric   <- combined_sc_min.py:116-164 (R over BOTH halves)
_trailing_average              <- combined_sc_min.py:167-180
_streaming_peak_detector       <- combined_sc_min.py:183-209
find_minn_peak (gated)         <- combined_sc_min.py:212-259
sc_gate_mask                   <- the inline gate construction of run_simulation, :337-351
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, metric_1d, out
from .core import N_FFT
from .minn import _trailing_average, minn_streaming_metric  # noqa: F401  (same functions in the reference)

SC_GATE_THRESHOLD = 0.6


def schmidl_cox_streaming_metric(rx, symbol_len: int = N_FFT):
    as_np = is_numpy_like(rx)
    arr = np.asarray(rx) if as_np else rx
    batched = arr.ndim == 3
    n = arr.shape[-1]
    if (symbol_len // 2 == 0 or symbol_len > n) and not batched:         # combined_sc_min.py:137-138
        if as_np:
            return np.zeros(0), np.zeros(0, dtype=np.complex128), np.zeros(0)
        z = torch.zeros(0, device=arr.device)
        return z, torch.zeros(0, dtype=torch.complex64, device=arr.device), z.clone()
    r = engine.metric(arr, "sc_both", int(symbol_len), want_pr=True, path="tile")
    sq = not batched
    return out(r.M, as_np, sq), out(r.P, as_np, sq), out(r.R, as_np, sq)


def sc_gate_mask(M_sc, threshold: float = SC_GATE_THRESHOLD):
    """gate = (M_sc / max(M_sc) >= threshold), seeded with argmax when empty (combined_sc_min.py:337-351)."""
    as_np = is_numpy_like(M_sc)
    return out(engine.sc_gate(metric_1d(M_sc), threshold), as_np)


def _streaming_peak_detector(metric, gate_mask):
    as_np = is_numpy_like(metric)
    g = np.asarray(gate_mask) if is_numpy_like(gate_mask) else gate_mask
    if g.shape[0] != (np.asarray(metric).shape[0] if as_np else metric.shape[0]):
        raise ValueError("gate_mask must match metric length")            # :185-186
    if (np.asarray(metric).size if as_np else metric.numel()) == 0:
        return None
    # smooth_win=1 on a non-negative metric is the identity: shift negatives up (argmax is shift-invariant)
    t = metric_1d(metric)
    t = t - torch.clamp(t.min(), max=0.0)
    pk = int(engine.find_minn_peak_gated(t, 1, g)[0].item())
    return None if pk < 0 else pk


def find_minn_peak(M, smooth_win: int = 8, gate_mask=None, search_bounds: tuple[int, int] | None = None) -> int:
    as_np = is_numpy_like(M)
    n = np.asarray(M).size if as_np else M.numel()
    if n == 0:
        return 0                                                           # :223-224
    if gate_mask is None:
        raise ValueError("Minn peak detection requires S&C gate mask")    # :228-229
    g = np.asarray(gate_mask) if is_numpy_like(gate_mask) else gate_mask
    if g.shape[0] != n:
        raise ValueError("gate_mask must match metric length")            # :230-231
    pk = int(engine.find_minn_peak_gated(metric_1d(M), smooth_win, g, search_bounds)[0].item())
    if pk == -3:
        raise ValueError("Minn peak detector received empty gate region")  # :245-246
    return pk
