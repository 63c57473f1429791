"""Synthetic frames for tests and bench.py (NOT on the hot path): preamble + pilot + data through a
measured CIR, AWGN and CFO -- the recipe of the reference's scripts (sc.py:181-202, minn.py:337-361,
channel.py:51-98, core.py:123-138), restated so that inputs can be generated on the GPU box where
/root/reference does not exist.  Host builders are numpy; `make_batch_device` fills a device batch
with torch (cuFFT convolution + Philox noise) -- input generation only, never timed.
"""
from __future__ import annotations

from pathlib import Path

import numpy as np

N_FFT, NUM_ACTIVE, CP, PRE_PAD, FS = 2048, 1200, 512, 1337, 30_720_000.0
_CIR_DATA = Path(__file__).resolve().parent / "data" / "channel_models.npz"


def centered_idx(width: int) -> np.ndarray:
    h = width // 2
    return np.concatenate((np.arange(-h, 0), np.arange(1, h + 1)))


def _to_time(idx: np.ndarray, vals: np.ndarray, n_fft: int = N_FFT) -> np.ndarray:
    spec = np.zeros(n_fft, dtype=complex)
    spec[(n_fft // 2 + idx) % n_fft] = vals
    td = np.fft.ifft(np.fft.ifftshift(spec))
    p = np.mean(np.abs(td) ** 2)
    return td / np.sqrt(p) if p > 0 else td


def _with_cp(sym: np.ndarray, cp: int = CP) -> np.ndarray:
    return np.concatenate((sym[-cp:], sym)) if cp > 0 else sym


def sc_preamble(rng: np.random.Generator) -> np.ndarray:
    """BPSK on the even active subcarriers -> two identical halves (sc.py:31-39)."""
    idx = centered_idx(NUM_ACTIVE)
    even = idx[idx % 2 == 0]
    return _with_cp(_to_time(even, rng.choice([-1.0, 1.0], size=even.size)))


def minn_preamble(rng: np.random.Generator) -> np.ndarray:
    """[A A -A -A] from a quarter-length BPSK symbol (structure of minn.py:30-56)."""
    q = N_FFT // 4
    idx = centered_idx(NUM_ACTIVE // 4)
    a = _to_time(idx, rng.choice([-1.0, 1.0], size=idx.size), q)
    sym = np.concatenate((a, a, -a, -a))
    return _with_cp(sym / np.sqrt(np.mean(np.abs(sym) ** 2)))


def qpsk_symbol(rng: np.random.Generator) -> np.ndarray:
    idx = centered_idx(NUM_ACTIVE)
    m = rng.integers(0, 4, size=idx.size)
    vals = (((m & 1) * 2 - 1) + 1j * (((m >> 1) & 1) * 2 - 1)) / np.sqrt(2.0)
    return _with_cp(_to_time(idx, vals))


def frame(rng: np.random.Generator, kind: str = "sc") -> np.ndarray:
    pre = sc_preamble(rng) if kind == "sc" else minn_preamble(rng)
    return np.concatenate((np.zeros(PRE_PAD, complex), pre, qpsk_symbol(rng), qpsk_symbol(rng)))


def load_cirs() -> dict:
    """cir1 / cir2 (2 RX channels x 1100 taps) from the package data file data/channel_models.npz (made from the reference's channel_models/*.csv by oracle/gen_golden.py)."""
    d = np.load(_CIR_DATA)
    return {"cir1": d["cir1"], "cir2": d["cir2"]}


def apply_channel_host(tx: np.ndarray, snr_db: float, rng: np.random.Generator, cir: np.ndarray | None, cfo_hz: float,
                       fs: float = FS) -> np.ndarray:
    """channel.apply_channel (one branch) + core.apply_cfo, numpy."""
    faded = tx if cir is None else np.convolve(tx, cir, mode="full")
    p = np.mean(np.abs(faded) ** 2)
    std = np.sqrt(p / (10 ** (snr_db / 10)) / 2)
    rx = faded + std * (rng.standard_normal(faded.shape) + 1j * rng.standard_normal(faded.shape))
    n = np.arange(rx.size, dtype=float)
    return rx * np.exp(1j * 2 * np.pi * cfo_hz * n / fs)


def tiled_stream_host(n_samples: int, seed: int, kind: str = "sc", cir: np.ndarray | None = None, snr_db: float = 10.0,
                      cfo_hz: float = 1000.0) -> np.ndarray:
    """One capture of n_samples complex64: frames tiled back to back, channel, AWGN, CFO (BASELINE cfg 2/3)."""
    rng = np.random.default_rng(seed)
    taps = 0 if cir is None else cir.size - 1
    fr = frame(rng, kind)
    reps = (n_samples - taps + fr.size - 1) // fr.size
    tx = np.tile(fr, reps)[: n_samples - taps]
    return apply_channel_host(tx, snr_db, rng, cir, cfo_hz).astype(np.complex64)


def aa_capture_host(n_samples: int, n_ant: int, seed: int, half_len: int = 512, snr_db: float = 10.0, cfo_hz: float = 500.0,
                    fs: float = 15_360_000.0, gap: int = 3000, int12: bool = False):
    """Antenna-array capture (n_ant, n_samples): [gap zeros][A][A] preambles (two identical halves of half_len samples,
    the structure of sync_aa.build_aa_preamble, sync_aa.py:699-711) repeated, a random phase and independent AWGN per
    antenna (sync_aa.py:591-634), CFO.  int12=True -> int16 IQ (n_ant, n_samples, 2) by the 12-bit ADC quantiser at
    full-scale ratio 2.0 (sync_aa.py:263-291)."""
    rng = np.random.default_rng(seed)
    half = (rng.choice([-1.0, 1.0], half_len) + 1j * rng.choice([-1.0, 1.0], half_len)) / np.sqrt(2.0)
    half = np.fft.ifft(np.fft.fft(half) * (np.abs(np.fft.fftfreq(half_len)) < 0.3))      # band-limit a little
    half /= np.sqrt(np.mean(np.abs(half) ** 2))
    unit = np.concatenate((np.zeros(gap, complex), half, half))
    tx = np.tile(unit, n_samples // unit.size + 1)[:n_samples]
    std = np.sqrt(1.0 / (10 ** (snr_db / 10)) / 2)
    t = np.arange(n_samples, dtype=float)
    rot = np.exp(1j * 2 * np.pi * cfo_hz * t / fs)
    out = np.empty((n_ant, n_samples), dtype=np.complex64)
    for a in range(n_ant):
        ph = np.exp(1j * rng.uniform(0, 2 * np.pi))
        out[a] = ((tx * ph + std * (rng.standard_normal(n_samples) + 1j * rng.standard_normal(n_samples))) * rot).astype(np.complex64)
    if not int12:
        return out
    peak = np.max(np.abs(np.concatenate((out.real, out.imag), axis=None)))
    q = np.clip(np.round(np.stack((out.real, out.imag), axis=-1) / (2.0 * peak / 2.0) * 2048.0 / 2.0), -2048, 2047)
    return q.astype(np.int16)


def make_batch_device(n_frames: int, n_samples: int, kind: str = "sc", seed: int = 0, device=None, n_base: int = 32,
                      chunk: int = 1024, want_iq: bool = False, full_scale_ratio: float = 4.0):
    """[n_frames, n_samples] complex64 on the device (and int16 IQ with want_iq): base frames tiled, cir1-ch1 (even) /
    cir2-ch1 (odd), SNR cycling {0,5,10,15,20} dB, CFO = linspace(-10 kHz, +10 kHz) (SURVEY.md 8d cfg 2), generated by the
    package's own impairment chain (ofs_channel_apply: overlap-save FIR, AWGN, CFO, ADC); torch only draws the Philox noise."""
    import torch
    from . import engine
    device = device or torch.device("cuda", torch.cuda.current_device())
    cirs = load_cirs()
    taps = cirs["cir1"].shape[1]
    n_tx = n_samples - (taps - 1)
    rng = np.random.default_rng(seed)
    base = np.stack([np.tile(frame(rng, kind), (n_tx + 9016) // 9017)[:n_tx] for _ in range(n_base)]).astype(np.complex64)
    base_d = torch.as_tensor(base).to(device)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    out = torch.empty((n_frames, n_samples), dtype=torch.complex64, device=device)
    iq = torch.empty((n_frames, n_samples, 2), dtype=torch.int16, device=device) if want_iq else None
    snrs = np.array([0.0, 5.0, 10.0, 15.0, 20.0])
    cfos = np.linspace(-10e3, 10e3, max(n_frames, 2))[:n_frames]
    for par, name in ((0, "cir1"), (1, "cir2")):
        idx = np.arange(par, n_frames, 2)
        for c0 in range(0, idx.size, chunk):
            f = idx[c0:c0 + chunk]
            noise = torch.view_as_complex(torch.randn((f.size, n_samples, 2), generator=gen, device=device, dtype=torch.float32))
            o, q = engine.channel_apply(base_d, cirs[name][1], row_of_stream=(f % n_base).astype(np.int32), unit_noise=noise,
                                        snr_db=snrs[f % 5], cfo_hz=cfos[f], fs=FS,
                                        full_scale=np.full(f.size, full_scale_ratio) if want_iq else None, want_iq=want_iq)
            fi = torch.as_tensor(f).to(device)
            out[fi] = o
            if want_iq:
                iq[fi] = q
    return (out, iq) if want_iq else out
