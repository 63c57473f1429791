"""Drop-in for the reference's zc_freq.py: make_pss_frequency_template (:54-59), compute_frequency_metric (:62-99)."""
from __future__ import annotations

import numpy as np

from . import engine
from ._shim import is_numpy_like, out
from .core import CYCLIC_PREFIX, N_FFT
from .zc import PSS_LENGTH, PSS_ROOT, build_pss_symbol, generate_zadoff_chu  # noqa: F401


def make_pss_frequency_template():
    half = PSS_LENGTH // 2
    bin_indices = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
    template_bins = generate_zadoff_chu(PSS_ROOT, PSS_LENGTH)
    return bin_indices, template_bins, float(np.sum(np.abs(template_bins) ** 2))


def compute_frequency_metric(rx_samples, bin_indices, template_bins, template_energy):
    as_np = is_numpy_like(rx_samples)
    arr = np.asarray(rx_samples) if as_np else rx_samples
    m = engine.zc_freq_metric(arr, bin_indices, template_bins, template_energy, N_FFT, CYCLIC_PREFIX)
    return out(m, as_np, squeeze=arr.ndim != 3)
