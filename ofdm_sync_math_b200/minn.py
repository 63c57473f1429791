"""Drop-in for the hot functions of the reference's minn.py.

minn_streaming_metric                <- minn.py:59-112
minn_streaming_metric_parameterized  <- minn.py:697-751
_trailing_average / find_minn_peak   <- minn.py:115-205
"""
from __future__ import annotations

import numpy as np
import torch

from . import engine
from ._shim import is_numpy_like, metric_1d, out
from .core import N_FFT


def minn_streaming_metric_parameterized(rx, symbol_len: int):
    as_np = is_numpy_like(rx)
    arr = np.asarray(rx) if as_np else rx
    batched = arr.ndim == 3
    n = arr.shape[-1]
    if max(n - symbol_len + 1, 0) <= 0 and not batched:     # minn.py:80-82, 723-724
        if as_np:
            return np.zeros(0), np.zeros(0, dtype=complex), np.zeros(0)
        z = torch.zeros(0, device=arr.device)
        return z, torch.zeros(0, dtype=torch.complex64, device=arr.device), z.clone()
    r = engine.metric(arr, "minn", int(symbol_len), want_pr=True, path="tile")
    sq = not batched
    return out(r.M, as_np, sq), out(r.P, as_np, sq), out(r.R, as_np, sq)


def minn_streaming_metric(rx):
    return minn_streaming_metric_parameterized(rx, N_FFT)


def _trailing_average(x, win: int):
    """minn.py:115-128 (computed by the find_minn_peak kernel with a wide-open gate)."""
    as_np = is_numpy_like(x)
    xa = np.asarray(x, dtype=float) if as_np else x
    if win <= 1:
        return xa.copy() if as_np else xa.clone()
    if (xa.size if as_np else xa.numel()) == 0:
        return xa.copy() if as_np else xa.clone()
    # the kernel smooths max(x, 0); shift so that negative inputs survive (average is affine)
    t = metric_1d(xa)
    lo = torch.clamp(t.min(), max=0.0)
    _, _, Ms = engine.find_minn_peak(t - lo + 1.0, smooth_win=win, gate_threshold=0.0, want_ms=True)
    return out(Ms + lo - 1.0, as_np)


def find_minn_peak(M, smooth_win: int = 8, gate_threshold: float = 0.5, search_bounds: tuple[int, int] | None = None):
    as_np = is_numpy_like(M)
    n = np.asarray(M).size if as_np else M.numel()
    if n == 0:
        raise ValueError("Minn metric is empty")                              # minn.py:142-143
    peak, span, Ms = engine.find_minn_peak(metric_1d(M), smooth_win, gate_threshold, search_bounds, want_ms=True)
    pk = int(peak[0].item())
    if pk == -2:
        raise ValueError("Minn metric did not produce a positive peak")       # minn.py:152-153
    s, e = (int(v) for v in span[0].tolist())
    if as_np:
        gate = np.zeros(n, dtype=bool)
        gate[s:e] = True
    else:
        gate = torch.zeros(n, dtype=torch.bool, device=M.device)
        gate[s:e] = True
    return pk, gate, out(Ms, as_np)


# ------------------------------------------------------------------------------------------------------------------
# the block-length sweep of minn.py on the batched engine.
SNR_DB = 0.0
CFO_HZ = 1000.0


def build_minn_preamble_parameterized(rng: np.random.Generator, symbol_len: int, cp_len: int, include_cp: bool = True) -> np.ndarray:
    """minn.build_minn_preamble_parameterized (minn.py:656-694): random BPSK quarter A -> [A A -A -A], unit power, CP."""
    if symbol_len % 4 != 0:
        raise ValueError(f"symbol_len must be divisible by 4, got {symbol_len}")
    a = rng.choice([-1.0, 1.0], size=symbol_len // 4) + 0j
    sym = np.concatenate((a, a, -a, -a))
    pw = np.mean(np.abs(sym) ** 2)
    if pw > 0:
        sym = sym / np.sqrt(pw)
    if not include_cp:
        return sym
    return np.concatenate((sym[-cp_len:] if cp_len <= symbol_len else sym, sym))


def compare_block_lengths(block_lengths, channel_name=None, snr_db=None):
    """minn.compare_block_lengths (minn.py:754-871): {N: {peak, par, pmr, timing_error, preamble_len, overhead_pct, metric,
    P_sum, R_sum}}.  Extension: snr_db may be a sequence -> {snr: {N: {...}}} with all SNRs of a block length evaluated as
    one batch on the device (what plot_snr_sweep, minn.py:1011-1022, loops over)."""
    from . import sweeps
    from .core import TX_PRE_PAD_SAMPLES
    many = isinstance(snr_db, (list, tuple, np.ndarray))
    snrs = [float(s) for s in snr_db] if many else [SNR_DB if snr_db is None else float(snr_db)]
    cir, delay = sweeps.channel_bank(channel_name)
    res = {s: {} for s in snrs}
    for N in block_lengths:
        rng = np.random.default_rng(0)
        cp_len = N // 4
        pre = build_minn_preamble_parameterized(rng, N, cp_len, include_cp=True)
        tx, frame_len = sweeps.two_frame_stream(pre, rng)
        base = dict(preamble_len=cp_len + N, overhead_pct=100.0 * (cp_len + N) / frame_len)
        rx = sweeps.received_batch(tx, snrs, cir, CFO_HZ)
        if rx.shape[2] - N + 1 <= 0:
            for s in snrs:
                res[s][N] = dict(peak=0, par=0, pmr=0, timing_error=0, **base)
            continue
        r = engine.metric(rx, "minn", int(N), want_pr=True, out_f64=True, path="tile")
        pk = sweeps.first_frame_argmax(r.M, TX_PRE_PAD_SAMPLES + frame_len + frame_len // 2)
        peak, par, pmr = (t.cpu().numpy() for t in sweeps.peak_statistics(r.M, pk))
        pk_h = pk.cpu().numpy()
        expected = TX_PRE_PAD_SAMPLES + delay + cp_len
        M, P, R = r.M.cpu().numpy(), r.P.cpu().numpy(), r.R.cpu().numpy()
        for i, s in enumerate(snrs):
            res[s][N] = dict(peak=float(peak[i]), par=float(par[i]), pmr=float(pmr[i]), timing_error=int(pk_h[i]) - expected,
                             metric=M[i], P_sum=P[i], R_sum=R[i], **base)
    return res if many else res[snrs[0]]


def run_block_length_comparison(channel_name=None) -> None:
    """minn.run_block_length_comparison (minn.py:874-896): the sweep over N = 256 ... 2048 as a printed table."""
    lengths = [256, 512, 1024, 2048]
    res = compare_block_lengths(lengths, channel_name)
    print(f"BLOCK LENGTH COMPARISON (Classical Minn) - {'Measured CIR ' + repr(channel_name) if channel_name else 'Flat AWGN'}")
    print(f"{'N':>6} | {'CP+N':>6} | {'Overhead':>8} | {'Peak':>10} | {'PAR':>8} | {'PMR':>8} | {'Timing':>8}")
    for N in lengths:
        r = res[N]
        print(f"{N:>6} | {r['preamble_len']:>6} | {r['overhead_pct']:>7.1f}% | {r['peak']:>10.4f} | {r['par']:>8.1f} | {r['pmr']:>8.2f} | {r['timing_error']:>+8}")
