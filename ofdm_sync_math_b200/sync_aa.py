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


# ------------------------------------------------------------------------------------------------------------------
# the grid sweep of sync_aa.py (run_single_test :669-823, run_grid_test :829-897) on the batched
# engine.  The transmit side (preamble / QPSK symbols, a 1024-point IFFT each) and the caller's seeded noise draws stay
# on the host in the reference's order; everything per received sample -- CIR convolution, AWGN scaling, CFO rotation,
# ADC quantiser, the detector's running sums and the gate FSM -- runs on the GPU, and all SNR x full-scale cases that
# share a (preamble length, channel) pair go through it as one batch: the default 135-case grid is 9 batches of 15.
# Plots are not produced (plot / plot_samples are accepted and ignored).
N_FFT = 1024
NUM_ACTIVE_SUBCARRIERS = 600
CYCLIC_PREFIX = 72
PREAMBLE_LENGTHS = [1024, 512, 256]
DEFAULT_PREAMBLE_LEN = 1024
ADC_BITS = 12
TX_PRE_PAD_SAMPLES = 500
TX_POST_PAD_SAMPLES = 500


def build_aa_preamble(total_length: int = DEFAULT_PREAMBLE_LEN):
    """sync_aa.build_aa_preamble (sync_aa.py:160-235): a Zadoff-Chu sequence on every K-th FFT bin of the active band
    (K = 2 N_FFT / total_length, DC skipped) -> time domain, first total_length samples = [A][A], unit power.
    -> (preamble, zc sequence, PAPR in dB)."""
    if total_length not in PREAMBLE_LENGTHS:
        raise ValueError(f"total_length must be one of {PREAMBLE_LENGTHS}, got {total_length}")
    step = 2 * N_FFT // total_length
    centre, reach = N_FFT // 2, NUM_ACTIVE_SUBCARRIERS // 2
    bins = np.arange(centre - reach, centre + reach + 1)
    bins = bins[(bins != centre) & (bins % step == 0)]
    count = bins.size
    root = 23 if count % 25 == 0 else 25
    k = np.arange(count)
    zc = np.exp(-1j * np.pi * root * k * (k + 1) / count)
    spec = np.zeros(N_FFT, dtype=complex)
    spec[bins] = zc
    td = (np.fft.ifft(spec) * np.sqrt(N_FFT))[:total_length]
    td = td / np.sqrt(np.mean(np.abs(td) ** 2))
    pw = np.abs(td) ** 2
    return td, zc, 10 * np.log10(np.max(pw) / np.mean(pw))


def build_random_qpsk_symbol(rng: np.random.Generator):
    """sync_aa.build_random_qpsk_symbol (sync_aa.py:238-257): one unit-power QPSK OFDM symbol with its cyclic prefix."""
    half = NUM_ACTIVE_SUBCARRIERS // 2
    idx = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
    q = rng.integers(0, 4, size=idx.size)
    vals = np.exp(1j * np.pi / 4 * (2 * q + 1)) / np.sqrt(2)
    spec = np.zeros(N_FFT, dtype=complex)
    spec[(N_FFT // 2 + idx) % N_FFT] = vals
    sym = np.fft.ifft(np.fft.ifftshift(spec)) * np.sqrt(N_FFT)
    sym = sym / np.sqrt(np.mean(np.abs(sym) ** 2))
    return np.concatenate((sym[-CYCLIC_PREFIX:], sym)), vals


def quantize_adc(samples, full_scale: float, bits: int = ADC_BITS) -> np.ndarray:
    """sync_aa.quantize_adc (sync_aa.py:263-291) on the device (ofs_channel_apply's quantiser stage)."""
    a = np.asarray(samples, dtype=np.complex128)
    o, _ = engine.channel_apply(a.reshape(1, -1), None, full_scale=float(full_scale), bits=bits)
    return o.cpu().numpy().reshape(a.shape)


def apply_cfo(samples, cfo_hz: float, sample_rate: float) -> np.ndarray:
    """sync_aa.apply_cfo (sync_aa.py:637-645), antennas on axis 0, on the device."""
    a = np.asarray(samples, dtype=np.complex128)
    o, _ = engine.channel_apply(np.atleast_2d(a), None, cfo_hz=float(cfo_hz), fs=float(sample_rate))
    return o.cpu().numpy().reshape(a.shape)


@dataclass
class TestResult:
    """sync_aa.TestResult (sync_aa.py:650-666)."""
    __test__ = False
    snr_db: float
    channel: str
    full_scale_ratio: float
    preamble_length: int
    timing_error: int
    cfo_applied_hz: float
    cfo_estimated_hz: float
    cfo_error_hz: float
    detected: bool
    num_events: int
    clipping_pct: float
    effective_bits: float
    metric_peak: float


def _grid_batch(preamble_length: int, channel_name, snr_values, full_scale_ratios, cfo_hz: float, seed: int,
                num_rx_antennas: int = 2) -> list[TestResult]:
    """All (snr, full-scale) cases of one (preamble length, channel) pair: sync_aa.py:694-823 for each, as one batch."""
    from .channel import load_measured_cir
    rng = np.random.default_rng(seed)
    half_len = preamble_length // 2
    preamble, _, _ = build_aa_preamble(preamble_length)
    pilot, _ = build_random_qpsk_symbol(rng)
    data, _ = build_random_qpsk_symbol(rng)
    tx = np.concatenate((np.zeros(TX_PRE_PAD_SAMPLES, complex), preamble, pilot, data, np.zeros(TX_POST_PAD_SAMPLES, complex)))
    A = num_rx_antennas
    if channel_name is None:
        taps, peak_offset = [None] * A, 0
    else:
        bank = load_measured_cir(channel_name)
        if bank.shape[0] < A:                                                   # sync_aa.py:611-614
            bank = np.tile(bank, (A // bank.shape[0] + 1, 1))
        bank = bank[:A]
        taps = list(bank)
        peak_offset = int(np.argmax(np.sum(np.abs(bank) ** 2, axis=0)))
    n_out = tx.size + (0 if channel_name is None else bank.shape[1] - 1)
    S, F = len(snr_values), len(full_scale_ratios)
    # the reference reseeds per case, so every case of this batch sees the same unit-variance draws (antenna 0 first)
    unit = [rng.standard_normal(n_out) + 1j * rng.standard_normal(n_out) for _ in range(A)]
    snr = np.asarray(snr_values, dtype=np.float64)
    rx = None
    for a in range(A):
        o, _ = engine.channel_apply(tx, taps[a], row_of_stream=np.zeros(S, np.int32), unit_noise=np.broadcast_to(unit[a], (S, n_out)),
                                    snr_db=snr, cfo_hz=float(cfo_hz), fs=SAMPLE_RATE_HZ)
        if rx is None:
            rx = torch.empty((S, A, n_out), dtype=o.dtype, device=o.device)
        rx[:, a] = o
    power = (rx.real ** 2 + rx.imag ** 2).mean(dim=(1, 2))
    rms = torch.sqrt(power)                                                     # sync_aa.py:727
    ratios = torch.as_tensor(np.asarray(full_scale_ratios, dtype=np.float64), device=rx.device)
    full_scale = rms[:, None] * ratios[None, :]                                 # [S, F]
    ros = (np.arange(S)[:, None, None] * A + np.arange(A)[None, None, :] + np.zeros((1, F, 1), int)).reshape(-1)
    fs_stream = full_scale[:, :, None].expand(S, F, A).reshape(-1).cpu().numpy()
    q, _ = engine.channel_apply(rx.reshape(S * A, n_out), None, row_of_stream=ros.astype(np.int32), full_scale=fs_stream, bits=ADC_BITS)
    q = q.reshape(S * F, A, n_out)
    # clipping statistics of the unquantised capture (sync_aa.py:294-315)
    lim = full_scale[:, :, None, None]
    over_re, over_im = rx.real.abs()[:, None] >= lim, rx.imag.abs()[:, None] >= lim
    clip_pct = (100.0 * (over_re | over_im).sum(dim=(2, 3)).to(torch.float64) / (A * n_out)).cpu().numpy()
    eff_bits = torch.clamp(ADC_BITS + torch.log2(rms[:, None] / full_scale), min=0.0).cpu().numpy()
    P, _, M, _ = engine.aa_metric_reference(q, half_len)
    events = engine.aa_events(M, P, half_len, DETECT_THRESHOLD, DETECT_HYSTERESIS, SAMPLE_RATE_HZ)
    m_max = M.max(dim=1).values.cpu().numpy() if n_out > half_len else np.zeros(S * F)
    true_start = TX_PRE_PAD_SAMPLES + peak_offset
    name = channel_name if channel_name else "awgn"
    res = []
    for s in range(S):
        for f in range(F):
            ev = events[s * F + f]
            common = dict(snr_db=snr_values[s], channel=name, full_scale_ratio=full_scale_ratios[f], preamble_length=preamble_length,
                          cfo_applied_hz=cfo_hz, clipping_pct=float(clip_pct[s, f]), effective_bits=float(eff_bits[s, f]))
            if len(ev):
                b = ev[int(np.argmax(ev["value"]))]                            # strongest event, first on ties (sync_aa.py:741)
                res.append(TestResult(timing_error=int(b["aux"]) - true_start, cfo_estimated_hz=float(b["cfo"]),
                                      cfo_error_hz=float(b["cfo"]) - cfo_hz, detected=True, num_events=len(ev),
                                      metric_peak=float(b["value"]), **common))
            else:
                res.append(TestResult(timing_error=0, cfo_estimated_hz=0.0, cfo_error_hz=cfo_hz, detected=False, num_events=0,
                                      metric_peak=float(m_max[s * F + f]), **common))
    return res


def run_single_test(snr_db: float, channel_name, full_scale_ratio: float, preamble_length: int = DEFAULT_PREAMBLE_LEN,
                    cfo_hz: float = 500.0, seed: int = 42, plot: bool = False, plot_dir=None) -> TestResult:
    """sync_aa.run_single_test (sync_aa.py:669-823): a batch of one."""
    return _grid_batch(preamble_length, channel_name, [snr_db], [full_scale_ratio], cfo_hz, seed)[0]


def run_grid_test(snr_values=(-5, 0, 5, 10, 15), channels=(None, "cir1", "cir2"), full_scale_ratios=(0.25, 0.5, 1.0, 1.5, 2.0),
                  preamble_lengths=tuple(PREAMBLE_LENGTHS), cfo_hz: float = 500.0, plot_samples: bool = True, seed: int = 42,
                  verbose: bool = False) -> list[TestResult]:
    """sync_aa.run_grid_test (sync_aa.py:829-897): results in the reference's order (preamble length, channel, SNR,
    full-scale ratio), one batched device pass per (preamble length, channel)."""
    out: list[TestResult] = []
    for plen in preamble_lengths:
        for ch in channels:
            out.extend(_grid_batch(plen, ch, list(snr_values), list(full_scale_ratios), cfo_hz, seed))
    if verbose:
        for i, r in enumerate(out):
            print(f"[{i + 1:3d}/{len(out)}] L={r.preamble_length // 2:3d} {r.channel:6s} SNR={r.snr_db:+3.0f}dB FS={r.full_scale_ratio:.2f}x -> "
                  f"{'hit ' if r.detected else 'miss'} timing_err={r.timing_error:+4d} cfo_err={r.cfo_error_hz:+7.1f}Hz "
                  f"clip={r.clipping_pct:5.1f}%")
    return out


def print_summary_table(results) -> None:
    """sync_aa.print_summary_table (sync_aa.py:900-960): timing error (or MISS) per SNR x full-scale ratio for every preamble
    length and channel, then the detection rate per (preamble length, channel)."""
    lengths = sorted({r.preamble_length for r in results}, reverse=True)
    chans = sorted({r.channel for r in results})
    snrs = sorted({r.snr_db for r in results})
    ratios = sorted({r.full_scale_ratio for r in results})
    key = {(r.preamble_length, r.channel, r.snr_db, r.full_scale_ratio): r for r in results}
    for plen in lengths:
        print(f"\nPREAMBLE LENGTH: {plen} samples (L={plen // 2})")
        for ch in chans:
            print(f"\n--- {ch.upper()} ---")
            print(f"{'SNR':>6s}" + "".join(f" | FS={fs:.2f}" for fs in ratios))
            for snr in snrs:
                cells = []
                for fs in ratios:
                    r = key.get((plen, ch, snr, fs))
                    cells.append("   N/A" if r is None else (f"{r.timing_error:+6d}" if r.detected else "  MISS"))
                print(f"{snr:+5.0f}dB" + "".join(f" | {c}" for c in cells))
    print("\nDETECTION RATE BY PREAMBLE LENGTH AND CHANNEL")
    for plen in lengths:
        print(f"\nPreamble L={plen // 2}:")
        for ch in chans:
            sel = [r for r in results if r.channel == ch and r.preamble_length == plen]
            hit = sum(r.detected for r in sel)
            print(f"  {ch:6s}: {hit}/{len(sel)} ({100 * hit / len(sel) if sel else 0:.0f}%)")


def main() -> None:
    """sync_aa.main (sync_aa.py:1073-1119) without the figures: preamble characteristics, the 135-case grid, the summary."""
    for plen in PREAMBLE_LENGTHS:
        pre, _, papr = build_aa_preamble(plen)
        a, b = pre[: plen // 2], pre[plen // 2:]
        corr = np.abs(np.vdot(a, b)) / (np.linalg.norm(a) * np.linalg.norm(b))
        print(f"  Length {plen:4d}: L={plen // 2:3d}, PAPR={papr:.2f}dB, duration={plen / SAMPLE_RATE_HZ * 1e6:.1f}us, [A][A] corr={corr:.6f}")
    results = run_grid_test(snr_values=[-5, 0, 5, 10, 15], channels=[None, "cir1", "cir2"], full_scale_ratios=[0.5, 1.0, 2.0],
                            preamble_lengths=PREAMBLE_LENGTHS, cfo_hz=500.0, plot_samples=False, verbose=True)
    print_summary_table(results)


if __name__ == "__main__":
    main()
