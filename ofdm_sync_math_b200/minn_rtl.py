"""Drop-in for the RTL-style Minn detector of the reference's minn_rtl.py (+ the integer datapath of ref/*.sv).

minn_rtl_streaming_metric  <- minn_rtl.py:667-733 (with _DelayLine/_RunningSum/_antenna_path :512-652)
detect_minn_rtl            <- minn_rtl.py:750-825
minn_rtl_int_metric        <- ref/minn_antenna_path.sv:63-194, ref/minn_preamble_detector.sv:247-325 (int64, floor-shift smoother)
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, out
from .core import N_FFT


@dataclass
class MinnRTLMetricState:
    corr_total: np.ndarray
    corr_positive: np.ndarray
    smooth_metric: np.ndarray
    energy_total: np.ndarray
    corr_scaled: np.ndarray
    energy_scaled: np.ndarray
    metric_valid: np.ndarray
    above_threshold: np.ndarray


@dataclass
class MinnRTLEvent:
    peak_index: int
    detected_index: int
    gate_segment: tuple


@dataclass
class MinnRTLDetection:
    events: list
    gate_mask: np.ndarray
    gate_segments: list


def minn_rtl_streaming_metric(rx, *, smooth_shift: int, threshold_value: int, threshold_frac_bits: int,
                              quarter_len: int | None = None) -> MinnRTLMetricState:
    as_np = is_numpy_like(rx)
    arr = np.asarray(rx, dtype=np.complex128) if as_np else rx
    if arr.ndim == 1:
        arr = arr[None, :]
    if quarter_len is None:
        quarter_len = N_FFT // 4
    if quarter_len <= 0:
        raise ValueError("quarter_len must be positive.")                      # minn_rtl.py:686-687
    d = engine.minn_rtl_metric(arr, quarter_len, smooth_shift, threshold_value, threshold_frac_bits)
    return MinnRTLMetricState(**{k: out(v, as_np) for k, v in d.items()})


def minn_rtl_int_metric(iq, *, smooth_shift: int, threshold_value: int, threshold_frac_bits: int,
                        quarter_len: int | None = None, lag_extra: int = 0) -> MinnRTLMetricState:
    """iq: int16 (antennas, n, 2).  corr_scaled / energy_scaled are the integer compare operands."""
    as_np = is_numpy_like(iq)
    if quarter_len is None:
        quarter_len = N_FFT // 4
    arr = np.asarray(iq) if as_np else iq
    if arr.ndim == 2:
        arr = arr[None]
    d = engine.minn_rtl_int(arr, quarter_len, smooth_shift, threshold_value, threshold_frac_bits, lag_extra)
    o = {k: out(v, as_np) for k, v in d.items()}
    o["corr_scaled"] = o["smooth_metric"] * (1 << threshold_frac_bits)
    o["energy_scaled"] = o["energy_total"] * int(threshold_value)
    return MinnRTLMetricState(**o)


def detect_minn_rtl(state: MinnRTLMetricState, *, hysteresis: int, timing_offset: int) -> MinnRTLDetection:
    as_np = is_numpy_like(state.corr_positive)
    t = lambda a: a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
    cp = t(state.corr_positive)
    n = cp.shape[-1]
    evs = engine.minn_rtl_events(cp, t(state.metric_valid), t(state.above_threshold), hysteresis, timing_offset)[0]
    events, segments = [], []
    for e in evs:
        seg = (int(e["gate_start"]), int(e["gate_end"]))
        segments.append(seg)
        if e["closed"]:                                                        # an open tail yields a segment, no event (:814-815)
            events.append(MinnRTLEvent(peak_index=int(e["peak_index"]), detected_index=int(e["aux"]), gate_segment=seg))
    gate_mask = np.zeros(n, dtype=bool)
    for s, e_ in segments:
        gate_mask[s:e_] = True
    if not as_np:
        gate_mask = torch.as_tensor(gate_mask, device=cp.device)
    return MinnRTLDetection(events=events, gate_mask=gate_mask, gate_segments=segments)
