"""Drop-in for the reference's channel.py: same names, arguments and return shapes (channel.py:18-98), CUDA underneath.
The AWGN is drawn on the host from the caller's numpy Generator in the reference's order (real part first, then the imaginary
part, channel.py:62-64), so a seeded run reproduces the reference sample for sample; the convolution, the power estimate,
the scaling and the sum run on the GPU (ofs_channel_apply)."""
from __future__ import annotations

import numpy as np

from . import engine, synth


def load_measured_cir(name: str) -> np.ndarray:
    """channel.load_measured_cir: (channels, taps) complex128.  The measured profiles ship as package data (data/channel_models.npz,
    made from the reference's channel_models/*.csv by oracle/gen_golden.py)."""
    cirs = synth.load_cirs()
    if name not in cirs:
        raise ValueError(f"Unknown channel profile '{name}'")
    return np.asarray(cirs[name], dtype=np.complex128)


def apply_channel(signal, snr_db: float, rng: np.random.Generator, channel_impulse_response=None) -> np.ndarray:
    """channel.apply_channel (channel.py:78-98) -> (branches, L) complex128."""
    signal = np.asarray(signal)
    if channel_impulse_response is None:
        rows = [(None, signal)]
    else:
        cir = np.asarray(channel_impulse_response)
        if cir.ndim == 1:
            cir = cir[np.newaxis, :]
        rows = [(taps, signal) for taps in cir]
    n_out = signal.size + (0 if channel_impulse_response is None else rows[0][0].size - 1)
    shape = (len(rows), n_out)
    unit = rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    out = np.empty(shape, dtype=np.complex128)
    for b, (taps, sig) in enumerate(rows):
        o, _ = engine.channel_apply(sig.astype(np.complex128), taps, unit_noise=unit[b:b + 1], snr_db=snr_db)
        out[b] = o.cpu().numpy()[0]
    return out
