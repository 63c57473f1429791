"""Drop-in for the Zadoff-Chu matched-filter detector that the reference's zc.py runs inline in
run_simulation (zc.py:105-130); lifted here into `zc_correlate` / `zc_detect`.

generate_zadoff_chu / build_pss_symbol (zc.py:34-46) are tiny host-side template builders (run once).
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, out
from .core import CYCLIC_PREFIX, N_FFT

PSS_LENGTH = 62
PSS_ROOT = 25


def generate_zadoff_chu(root: int, length: int) -> np.ndarray:
    n = np.arange(length)
    return np.exp(-1j * np.pi * root * n * (n + 1) / length)


def build_pss_symbol(include_cp: bool = True, root: int = PSS_ROOT) -> np.ndarray:
    """PSS time-domain symbol: 62 ZC tones centred on DC (skipped), unit power (zc.py:39-46, core.py:13-44)."""
    half = PSS_LENGTH // 2
    idx = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
    spectrum = np.zeros(N_FFT, dtype=complex)
    spectrum[(N_FFT // 2 + idx) % N_FFT] = generate_zadoff_chu(root, PSS_LENGTH)
    td = np.fft.ifft(np.fft.ifftshift(spectrum))
    td = td / np.sqrt(np.mean(np.abs(td) ** 2))
    return np.concatenate((td[-CYCLIC_PREFIX:], td)) if include_cp else td


def zc_correlate(rx_samples, pss_reference=None):
    """combined_corr of zc.py:105-126: branch-summed matched filter, normalised after the sum.
    Returns the complex correlation of length L + len(ref) - 1."""
    as_np = is_numpy_like(rx_samples)
    ref = build_pss_symbol(include_cp=False) if pss_reference is None else np.asarray(pss_reference)
    arr = np.asarray(rx_samples) if as_np else rx_samples
    corr, _ = engine.zc_matched_filter(arr, ref, mode=0)
    return out(corr, as_np, squeeze=arr.ndim != 3)


def zc_detect(rx_samples, pss_reference=None):
    """(combined_corr, peak_index, detected_start) of zc.py:105-130."""
    as_np = is_numpy_like(rx_samples)
    ref = build_pss_symbol(include_cp=False) if pss_reference is None else np.asarray(pss_reference)
    arr = np.asarray(rx_samples) if as_np else rx_samples
    corr, mag = engine.zc_matched_filter(arr, ref, mode=0)
    peak = engine.argmax(mag)
    if arr.ndim == 3:
        return out(corr, as_np, False), out(peak, as_np, False), out(torch.clamp(peak - len(ref) + 1, min=0), as_np, False)
    pk = int(peak[0].item())
    return out(corr, as_np), pk, max(pk - len(ref) + 1, 0)
