"""System constants of the reference (core.py:6-10, sync_aa.py:99-122) -- the defaults of the shims."""
N_FFT = 2048
NUM_ACTIVE_SUBCARRIERS = 1200
CYCLIC_PREFIX = 512
TX_PRE_PAD_SAMPLES = 1337
SAMPLE_RATE_HZ = 30_720_000.0

AA_N_FFT = 1024
AA_SAMPLE_RATE_HZ = 15_360_000.0
AA_PREAMBLE_HALF_LEN = 512
AA_DETECT_THRESHOLD = 0.15
AA_DETECT_HYSTERESIS = 128


# ---- functions: same signatures as the reference's core.py, CUDA underneath -------------------------
def _np():
    import numpy as np
    return np


def apply_cfo(samples, cfo_hz: float, fs_hz: float):
    """core.apply_cfo (core.py:123-138): 1-D or 2-D (branches, L) -> same shape, complex128."""
    np = _np()
    from . import engine
    x = np.asarray(samples)
    if x.ndim not in (1, 2):
        raise ValueError("samples must be 1D or 2D")
    out, _ = engine.channel_apply(np.atleast_2d(x).astype(np.complex128), None, cfo_hz=cfo_hz, fs=fs_hz)
    out = out.cpu().numpy()
    return out[0] if x.ndim == 1 else out


def _cfo(rx, start, n_fft, cp_len, fs_hz, mode, span=None, win_len=None):
    np = _np()
    from . import engine
    x = np.asarray(rx)
    if x.ndim == 1:
        x = x[np.newaxis, :]
    cfo, bd, _ = engine.cp_cfo(x[np.newaxis].astype(np.complex128), int(start), n_fft, cp_len, fs_hz, mode, span, win_len)
    c = float(cfo.cpu().numpy()[0])
    if c != c:
        raise ValueError("operands could not be broadcast together: the CP windows leave the capture")
    return c, int(bd.cpu().numpy()[0])


def estimate_cfo_from_cp(rx, start: int, n_fft: int, cp_len: int, fs_hz: float) -> float:
    """core.py:179-196."""
    return _cfo(rx, start, n_fft, cp_len, fs_hz, "plain")[0]


def estimate_cfo_from_cp_robust(rx, cp_start_est: int, n_fft: int, cp_len: int, fs_hz: float, span=None, win_len=None) -> float:
    """core.py:199-230."""
    return _cfo(rx, cp_start_est, n_fft, cp_len, fs_hz, "robust", span, win_len)[0]


def estimate_cfo_from_cp_peak(rx, cp_start_est: int, n_fft: int, cp_len: int, fs_hz: float, span=None) -> float:
    """core.py:233-265."""
    return _cfo(rx, cp_start_est, n_fft, cp_len, fs_hz, "peak", span)[0]


def estimate_cfo_from_cp_peak_with_index(rx, cp_start_est: int, n_fft: int, cp_len: int, fs_hz: float, span=None):
    """core.py:268-308 -> (cfo_hz, best_d)."""
    return _cfo(rx, cp_start_est, n_fft, cp_len, fs_hz, "peak", span)


def find_cp_start_via_corr(rx, est_start: int, n_fft: int, cp_len: int, search_half: int = 1024) -> int:
    """core.py:311-336.  An empty search range returns est_start without touching the data (core.py:323-324)."""
    np = _np()
    x = np.asarray(rx)
    L = x.shape[-1]
    lo = max(0, int(est_start) - int(search_half))
    hi = min(L - (n_fft + cp_len), int(est_start) + int(search_half))
    if hi <= lo:
        return int(est_start)
    return _cfo(rx, est_start, n_fft, cp_len, 1.0, "peak", search_half)[1]
