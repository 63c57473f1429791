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


# ------------------------------------------------------------------------------------------------------------------
# the segment-length sweep of minn_rtl.py on the batched engine.
SNR_DB = 0.0
CFO_HZ = 1000.0
SMOOTH_SHIFT = 3
THRESH_FRAC_BITS = 15
THRESH_VALUE = int(0.10 * (1 << THRESH_FRAC_BITS))
HYSTERESIS = 2
TIMING_OFFSET = 0
PREAMBLE_SEQ_TYPE = "qpsk_freq"


def _generate_zc_sequence(length: int, root: int = 1) -> np.ndarray:
    """minn_rtl.py:206-228."""
    n = np.arange(length)
    return np.exp(-1j * np.pi * root * n * ((n + 1) if length % 2 else n) / length)


def _quarter_comb(values_for) -> np.ndarray:
    """Time-domain symbol whose active subcarriers sit on every 4th bin (minn_rtl.py:255-262 and siblings)."""
    from .core import NUM_ACTIVE_SUBCARRIERS
    from .synth import centered_idx
    idx = centered_idx(NUM_ACTIVE_SUBCARRIERS)
    idx = idx[idx % 4 == 0]
    spec = np.zeros(N_FFT, dtype=complex)
    spec[(N_FFT // 2 + idx) % N_FFT] = values_for(idx.size)
    return np.fft.ifft(np.fft.ifftshift(spec))


def _generate_base_sequence(seq_type: str, length: int, rng: np.random.Generator | None = None) -> np.ndarray:
    """minn_rtl._generate_base_sequence (minn_rtl.py:231-332): the quarter sequence A, unit power."""
    Q = length
    needs_rng = {"bpsk_freq", "qpsk_freq", "random_phase"}
    if seq_type in needs_rng and rng is None:
        raise ValueError(f"rng required for {seq_type}")
    if seq_type == "bpsk_freq":
        a = _quarter_comb(lambda k: rng.choice([-1.0, 1.0], size=k))[:Q]
    elif seq_type == "qpsk_freq":
        a = _quarter_comb(lambda k: np.exp(1j * np.pi / 4 * (2 * rng.choice([0, 1, 2, 3], size=k) + 1)))[:Q]
    elif seq_type == "zc_time":
        a = _generate_zc_sequence(Q, root=7)
    elif seq_type == "zc_freq":
        a = _quarter_comb(lambda k: np.exp(-1j * np.pi * 7 * np.arange(k) * np.arange(k) / k))[:Q]
    elif seq_type == "chirp":
        n = np.arange(Q)
        a = np.exp(1j * np.pi * n * n / Q)
    elif seq_type == "gold":
        s1, s2 = 0b1010101010, 0b1100110011                                    # two 10-bit LFSRs, outputs XORed (:296-309)
        bits = np.zeros(Q, dtype=int)
        for i in range(Q):
            bits[i] = ((s1 >> 9) ^ (s2 >> 9)) & 1
            f1 = ((s1 >> 9) ^ (s1 >> 6)) & 1
            f2 = ((s2 >> 9) ^ (s2 >> 8) ^ (s2 >> 5) ^ (s2 >> 3)) & 1
            s1, s2 = ((s1 << 1) | f1) & 0x3FF, ((s2 << 1) | f2) & 0x3FF
        a = 2.0 * bits - 1.0 + 0j
    elif seq_type == "const":
        a = np.ones(Q, dtype=complex)
    elif seq_type == "random_phase":
        a = np.exp(1j * rng.uniform(0, 2 * np.pi, Q))
    else:
        raise ValueError(f"Unknown sequence type: {seq_type}")
    pw = np.mean(np.abs(a) ** 2)
    return a / np.sqrt(pw) if pw > 0 else a


def build_minn_preamble_generic(seq_type: str, rng: np.random.Generator | None = None, Q: int | None = None) -> np.ndarray:
    """minn_rtl.build_minn_preamble_generic (minn_rtl.py:335-358): [-A +A +A -A -A], 5 Q samples, unit power."""
    a = _generate_base_sequence(seq_type, N_FFT // 4 if Q is None else Q, rng)
    pre = np.concatenate((-a, a, a, -a, -a))
    pw = np.mean(np.abs(pre) ** 2)
    return pre / np.sqrt(pw) if pw > 0 else pre


def compare_q_values(q_values, channel_name=None, snr_db=None):
    """minn_rtl.compare_q_values (minn_rtl.py:1493-1592): {Q: {peak, par, pmr, timing_error, preamble_len, overhead_pct}}.
    Extension: snr_db (default SNR_DB) may be a sequence -> {snr: {Q: {...}}}, all SNRs of a Q as one device batch."""
    from . import sweeps
    from .core import TX_PRE_PAD_SAMPLES
    many = isinstance(snr_db, (list, tuple, np.ndarray))
    snrs = [float(s) for s in snr_db] if many else [SNR_DB if snr_db is None else float(snr_db)]
    cir, delay = sweeps.channel_bank(channel_name)
    res = {s: {} for s in snrs}
    for Q in q_values:
        rng = np.random.default_rng(0)
        pre = build_minn_preamble_generic(PREAMBLE_SEQ_TYPE, rng, Q=Q)
        tx, frame_len = sweeps.two_frame_stream(pre, rng)
        rx = sweeps.received_batch(tx, snrs, cir, CFO_HZ)
        d = engine.minn_rtl_metric(rx, Q, SMOOTH_SHIFT, THRESH_VALUE, THRESH_FRAC_BITS)
        evs = engine.minn_rtl_events(d["corr_positive"], d["metric_valid"], d["above_threshold"], HYSTERESIS, TIMING_OFFSET)
        fallback = engine.argmax(d["smooth_metric"]).cpu().numpy()               # no event: arg-max of the smoothed metric (:1556-1558)
        target = sweeps.pilot_start(TX_PRE_PAD_SAMPLES, delay, 5 * Q)
        pk, terr = np.empty(len(snrs), np.int64), np.empty(len(snrs), np.int64)
        for i in range(len(snrs)):
            closed = [e for e in evs[i] if e["closed"]]
            if closed:
                pk[i], terr[i] = int(closed[0]["peak_index"]), int(closed[0]["aux"]) - target
            else:
                pk[i], terr[i] = int(fallback[i]), int(fallback[i]) - target
        peak, par, pmr = (t.cpu().numpy() for t in sweeps.peak_statistics(d["corr_positive"], torch.as_tensor(pk)))
        for i, s in enumerate(snrs):
            res[s][Q] = dict(peak=float(peak[i]), par=float(par[i]), pmr=float(pmr[i]), timing_error=int(terr[i]),
                             preamble_len=5 * Q, overhead_pct=100.0 * 5 * Q / frame_len)
    return res if many else res[snrs[0]]


def run_q_comparison(channel_name=None) -> None:
    """minn_rtl.run_q_comparison (minn_rtl.py:1595-1617): the sweep over Q = 64 ... 512 as a printed table."""
    qs = [64, 128, 256, 512]
    res = compare_q_values(qs, channel_name)
    print(f"Q VALUE COMPARISON - {'Measured CIR ' + repr(channel_name) if channel_name else 'Flat AWGN'}")
    print(f"{'Q':>6} | {'5Q':>6} | {'Overhead':>8} | {'Peak':>12} | {'PAR':>8} | {'PMR':>8} | {'Timing':>8}")
    for q in qs:
        r = res[q]
        print(f"{q:>6} | {r['preamble_len']:>6} | {r['overhead_pct']:>7.1f}% | {r['peak']:>12.1f} | {r['par']:>8.1f} | {r['pmr']:>8.2f} | {r['timing_error']:>+8}")
