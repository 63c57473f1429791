"""Drop-in for the hot functions of the reference's zc_v2.py.

matched_filter_correlation  <- zc_v2.py:244-254      normalize_correlation  <- zc_v2.py:257-271
zc_streaming_detection      <- zc_v2.py:288-336      detect_zc_peaks        <- zc_v2.py:360-450
detect_zc_preamble          <- zc_v2.py:456-516
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, metric_1d, out
from .core import N_FFT
from .zc import PSS_LENGTH, PSS_ROOT, generate_zadoff_chu  # noqa: F401
from .zc import build_pss_symbol as _build_pss

CORR_WINDOW_SIZE = N_FFT
THRESH_FRAC_BITS = 15
THRESH_VALUE = int(4.0 * (1 << THRESH_FRAC_BITS) / CORR_WINDOW_SIZE)
MIN_CORR_MAG = 0.3
HYSTERESIS = 256


def build_pss_symbol(include_cp: bool = False) -> np.ndarray:
    return _build_pss(include_cp=include_cp)


@dataclass
class ZCDetectionState:
    corr_mag: np.ndarray
    local_sum: np.ndarray
    corr_scaled: np.ndarray
    thresh_scaled: np.ndarray
    above_threshold: np.ndarray
    metric_valid: np.ndarray


@dataclass
class ZCDetectionEvent:
    peak_index: int
    peak_value: float
    gate_start: int
    gate_end: int
    detected_start: int


@dataclass
class ZCDetectionResult:
    events: list
    gate_mask: np.ndarray
    state: ZCDetectionState


def matched_filter_correlation(rx_samples, reference):
    as_np = is_numpy_like(rx_samples)
    corr, _ = engine.zc_matched_filter(np.asarray(rx_samples) if as_np else rx_samples, reference, mode=2)
    return out(corr, as_np)


def normalize_correlation(corr, rx_samples, reference):
    """zc_v2.py:257-271: the SUPPLIED corr divided by ||ref|| * sqrt(max(sliding energy of rx_samples, 1e-12)) -- only the
    normaliser is computed from rx_samples (float64 prefix on the device, ofs_zc_normalize)."""
    as_np = is_numpy_like(rx_samples) and is_numpy_like(corr)
    c = engine.zc_normalize(np.asarray(corr) if is_numpy_like(corr) else corr,
                            np.asarray(rx_samples) if is_numpy_like(rx_samples) else rx_samples, reference)
    return out(c, as_np)


def zc_streaming_detection(corr_mag, window_size: int = CORR_WINDOW_SIZE, thresh_value: int = THRESH_VALUE,
                           thresh_frac_bits: int = THRESH_FRAC_BITS, min_corr_mag: float = MIN_CORR_MAG) -> ZCDetectionState:
    as_np = is_numpy_like(corr_mag)
    m = metric_1d(corr_mag)
    ls, valid, above = engine.zc_streaming_detection(m, window_size, thresh_value, thresh_frac_bits, min_corr_mag)
    cm = out(m, as_np)
    lsum = out(ls, as_np)
    return ZCDetectionState(corr_mag=cm, local_sum=lsum, corr_scaled=cm * float(1 << thresh_frac_bits),
                            thresh_scaled=lsum * float(thresh_value), above_threshold=out(above, as_np),
                            metric_valid=out(valid, as_np))


def detect_zc_peaks(state: ZCDetectionState, reference_length: int, hysteresis: int = HYSTERESIS) -> ZCDetectionResult:
    as_np = is_numpy_like(state.corr_mag)
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
    evs, gm = engine.zc_events(metric_1d(state.corr_mag), t(state.metric_valid), t(state.above_threshold), reference_length,
                               hysteresis)
    events = [ZCDetectionEvent(peak_index=int(e["peak_index"]), peak_value=float(e["value"]), gate_start=int(e["gate_start"]),
                               gate_end=int(e["gate_end"]), detected_start=int(e["aux"])) for e in evs[0]]
    return ZCDetectionResult(events=events, gate_mask=out(gm, as_np), state=state)


def detect_zc_preamble(rx_samples, window_size: int = CORR_WINDOW_SIZE, thresh_value: int = THRESH_VALUE,
                       thresh_frac_bits: int = THRESH_FRAC_BITS, min_corr_mag: float = MIN_CORR_MAG,
                       hysteresis: int = HYSTERESIS, normalize: bool = True) -> ZCDetectionResult:
    as_np = is_numpy_like(rx_samples)
    reference = build_pss_symbol(include_cp=False)
    arr = np.asarray(rx_samples) if as_np else rx_samples
    _, mag = engine.zc_matched_filter(arr, reference, mode=1 if normalize else 2)
    state = zc_streaming_detection(out(mag, as_np), window_size, thresh_value, thresh_frac_bits, min_corr_mag)
    return detect_zc_peaks(state, reference_length=len(reference), hysteresis=hysteresis)
