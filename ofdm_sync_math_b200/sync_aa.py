"""Drop-in for the [A][A] streaming detector of the reference's sync_aa.py.

aa_detect_streaming  <- sync_aa.py:421-571 (loop 1: per-antenna P, R, antenna combining, M; loop 2: gate FSM,
                        peak of |P|^2, CFO = angle(P) fs / (2 pi L), frame_start = peak - 2L + 1)
order="reference" (default) evaluates the reference's running-sum recurrences in its exact operation order
(bit-equal P; needed for the docs/detector_test_vector.csv tie); order="scan" uses the
parallel float64 prefix-sum kernel (same values to ~1e-13, the throughput path for many antennas).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, out
from .core import AA_DETECT_HYSTERESIS, AA_DETECT_THRESHOLD, AA_PREAMBLE_HALF_LEN, AA_SAMPLE_RATE_HZ

PREAMBLE_HALF_LEN = AA_PREAMBLE_HALF_LEN
DETECT_THRESHOLD = AA_DETECT_THRESHOLD
DETECT_HYSTERESIS = AA_DETECT_HYSTERESIS
SAMPLE_RATE_HZ = AA_SAMPLE_RATE_HZ


@dataclass
class AADetectorState:
    P: np.ndarray
    R: np.ndarray
    M: np.ndarray
    valid: np.ndarray


@dataclass
class AADetectionEvent:
    peak_index: int
    P_at_peak: complex
    M_at_peak: float
    gate_start: int
    gate_end: int
    cfo_hz: float
    frame_start: int


@dataclass
class AADetectionResult:
    events: list
    state: AADetectorState
    num_antennas: int


def aa_detect_streaming(rx_samples, L: int = PREAMBLE_HALF_LEN, threshold: float = DETECT_THRESHOLD,
                        hysteresis: int = DETECT_HYSTERESIS, sample_rate: float = SAMPLE_RATE_HZ,
                        order: str = "reference") -> AADetectionResult:
    as_np = is_numpy_like(rx_samples)
    arr = np.asarray(rx_samples) if as_np else rx_samples
    if arr.ndim == 1:
        arr = arr[None, :]
    if arr.ndim != 2:
        raise ValueError("rx_samples must be (num_antennas, num_samples) or (num_samples,)")
    na, n = arr.shape
    if order == "reference":
        P, R, M, valid = engine.aa_metric_reference(arr, L)
    else:
        r = engine.metric(arr, "aa", L, want_pr=True, out_f64=True, path="tile")
        P, R, M = r.P, r.R, r.M
        valid = (torch.arange(n, device=M.device) >= L).to(torch.uint8)[None]
    evs = engine.aa_events(M, P, L, threshold, hysteresis, sample_rate)[0]
    events = [AADetectionEvent(peak_index=int(e["peak_index"]), P_at_peak=complex(e["p_re"], e["p_im"]),
                               M_at_peak=float(e["value"]), gate_start=int(e["gate_start"]), gate_end=int(e["gate_end"]),
                               cfo_hz=float(e["cfo"]), frame_start=int(e["aux"])) for e in evs]
    state = AADetectorState(P=out(P, as_np), R=out(R, as_np), M=out(M, as_np), valid=out(valid, as_np))
    return AADetectionResult(events=events, state=state, num_antennas=int(na))
