"""Batched machinery for the parameter sweeps of the reference scripts:
minn.compare_block_lengths (minn.py:754-871), minn_rtl.compare_q_values (minn_rtl.py:1493-1592).

Each sweep point builds a two-frame transmit stream on the host (a handful of IFFTs), then everything per received sample
runs on the device: CIR convolution, AWGN at the requested SNR, CFO (ofs_channel_apply), the timing metric, the first-frame
arg-max and the noise-floor statistics.  A list of SNRs becomes the frames axis of one batch: the reference reseeds its
generators at every sweep point, so all SNRs of a point share the same unit-variance noise draws."""
from __future__ import annotations

import numpy as np
import torch

from . import engine, synth
from .core import CYCLIC_PREFIX, SAMPLE_RATE_HZ, TX_PRE_PAD_SAMPLES


def channel_bank(channel_name):
    """The (<= 2)-branch CIR the sweeps use (minn.py:777-780) and its strongest-path delay (core.py:113-120)."""
    if not channel_name:
        return None, 0
    from .channel import load_measured_cir
    bank = load_measured_cir(channel_name)
    cir = bank[:2].copy() if bank.shape[0] > 2 else bank.copy()
    agg = np.sum(np.abs(cir) ** 2, axis=0)
    return cir, (int(np.argmax(agg)) if np.any(agg) else 0)


def two_frame_stream(preamble: np.ndarray, rng: np.random.Generator):
    """[pre-pad][preamble pilot data][one frame of silence][preamble pilot data] (minn.py:795-810) -> (tx, frame_len)."""
    pilot = synth.qpsk_symbol(rng)
    data = synth.qpsk_symbol(rng)
    frame = np.concatenate((preamble, pilot, data))
    gap = np.zeros(frame.size, dtype=complex)
    return np.concatenate((np.zeros(TX_PRE_PAD_SAMPLES, dtype=complex), frame, gap, frame)), frame.size


def received_batch(tx: np.ndarray, snr_list, cir, cfo_hz: float, noise_seed: int = 0) -> torch.Tensor:
    """channel.apply_channel (channel.py:78-98, generator seeded noise_seed as at minn.py:813) + core.apply_cfo for every SNR
    of snr_list at once -> complex128 [len(snr_list), branches, n_out] on the device."""
    rng = np.random.default_rng(noise_seed)
    B = 1 if cir is None else cir.shape[0]
    n_out = tx.size + (0 if cir is None else cir.shape[1] - 1)
    unit = rng.standard_normal((B, n_out)) + 1j * rng.standard_normal((B, n_out))
    S = len(snr_list)
    snr = np.asarray(snr_list, dtype=np.float64)
    rx = None
    for b in range(B):
        o, _ = engine.channel_apply(tx.astype(np.complex128), None if cir is None else cir[b], row_of_stream=np.zeros(S, np.int32),
                                    unit_noise=np.broadcast_to(unit[b], (S, n_out)), snr_db=snr, cfo_hz=float(cfo_hz), fs=SAMPLE_RATE_HZ)
        if rx is None:
            rx = torch.empty((S, B, n_out), dtype=o.dtype, device=o.device)
        rx[:, b] = o
    return rx


def peak_statistics(metric: torch.Tensor, peak_idx: torch.Tensor, guard: int = 500, skip: int = TX_PRE_PAD_SAMPLES):
    """Peak value, peak / mean(noise) and peak / max(noise), the noise floor being the metric outside [peak - guard,
    peak + guard) and past the first `skip` samples (minn.py:841-858).  metric [S, n], peak_idx [S] -> three float64 [S]."""
    S, n = metric.shape
    pos = torch.arange(n, device=metric.device)[None, :]
    pk = peak_idx.to(metric.device)[:, None]
    keep = ((pos < pk - guard) | (pos >= pk + guard)) & (pos >= skip)
    m = metric.to(torch.float64)
    peak = m.gather(1, pk).squeeze(1)
    cnt = keep.sum(dim=1)
    mean = torch.where(keep, m, torch.zeros_like(m)).sum(dim=1) / cnt.clamp(min=1)
    mx = torch.where(keep, m, torch.full_like(m, -float("inf"))).max(dim=1).values
    inf = torch.full_like(peak, float("inf"))
    par = torch.where((cnt > 0) & (mean > 0), peak / mean, inf)
    pmr = torch.where((cnt > 0) & (mx > 0), peak / mx, inf)
    return peak, par, pmr


def first_frame_argmax(metric: torch.Tensor, end: int) -> torch.Tensor:
    """First maximum of metric[:, :end] per row (np.argmax at minn.py:836) with the library's arg-max kernel."""
    end = max(1, min(int(end), metric.shape[1]))
    return engine.argmax(metric[:, :end].contiguous())


def pilot_start(frame_start: int, delay: int, preamble_len: int) -> int:
    return frame_start + delay + preamble_len + CYCLIC_PREFIX
