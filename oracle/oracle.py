"""Python face of the CPU oracle (oracle/ofs_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this module.  It mirrors the reference's function signatures (amcolex/ofdm-sync-math)
so that parity tests read like calls to the reference; every function cites the reference
file:line it restates.  Parity status: pinned by tests/test_oracle_golden.py against
tests/golden/*.npz (outputs of the unmodified reference, made by oracle/gen_golden.py).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from dataclasses import dataclass
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libofs_oracle.so"
_lib = None

i64 = C.c_int64
_dp = C.POINTER(C.c_double)


def build(force: bool = False) -> Path:
    """Compile oracle/ofs_oracle.c with gcc (make)."""
    src = _HERE / "ofs_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-s", "-B"], check=True)
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        for name in ("orc_sc_metric", "orc_scb_metric", "orc_minn_metric", "orc_park_metric",
                     "orc_find_plateau_end", "orc_find_minn_peak", "orc_find_minn_peak_gated",
                     "orc_detect_zc_peaks", "orc_zc_freq_metric", "orc_aa_events", "orc_detect_minn_rtl",
                     "orc_detect_minn_rtl_int", "orc_metric_prefix_c64"):
            getattr(_lib, name).restype = i64
    return _lib


def _p(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _c128_2d(rx) -> np.ndarray:
    x = np.asarray(rx)
    if x.ndim == 1:
        x = x[np.newaxis, :]
    return np.ascontiguousarray(x, dtype=np.complex128)


def _cplx(a: np.ndarray) -> np.ndarray:
    return a.view(np.complex128).reshape(-1)


# ----------------------------------------------------------------------------- metrics
def _metric3(fn, rx, N):
    x = _c128_2d(rx)
    nb, L = x.shape
    n = max(L - N + 1, 0)
    M = np.zeros(n); P = np.zeros(2 * n); R = np.zeros(n)
    if n > 0:
        got = fn(_p(x), i64(nb), i64(L), i64(N), _p(M), _p(P), _p(R))
        if got == 0:
            return np.zeros(0), np.zeros(0, dtype=complex), np.zeros(0)
    return M, _cplx(P), R


def sc_streaming_metric(rx, n_fft: int = 2048):
    """sc.py:42-78."""
    return _metric3(lib().orc_sc_metric, rx, n_fft)


def schmidl_cox_streaming_metric(rx, symbol_len: int = 2048):
    """combined_sc_min.py:116-164 (R over both halves)."""
    return _metric3(lib().orc_scb_metric, rx, symbol_len)


def minn_streaming_metric(rx, symbol_len: int = 2048):
    """minn.py:59-112 / minn.py:697-751 / combined_sc_min.py:60-113."""
    return _metric3(lib().orc_minn_metric, rx, symbol_len)


minn_streaming_metric_parameterized = minn_streaming_metric


def park_streaming_metric(rx, n_fft: int = 2048):
    """park.py:64-114."""
    x = _c128_2d(rx)
    nb, L = x.shape
    n = max(L - 2 * (n_fft // 2), 0)
    ds = np.zeros(n, dtype=np.int64); M = np.zeros(n); P = np.zeros(2 * n); E = np.zeros(n)
    got = lib().orc_park_metric(_p(x), i64(nb), i64(L), i64(n_fft), _p(ds), _p(M), _p(P), _p(E)) if n else 0
    if got == 0:
        return np.zeros(0, dtype=int), np.zeros(0), np.zeros(0, dtype=complex), np.zeros(0)
    return ds, M, _cplx(P), E


# ----------------------------------------------------------------------------- detectors
def trailing_average(x, win: int) -> np.ndarray:
    """minn.py:115-128."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    y = np.empty_like(x)
    lib().orc_trailing_average(_p(x), i64(x.size), i64(win), _p(y))
    return y


def find_plateau_end_from_metric(M, cp_len: int, lookahead: int | None = None, smooth_win: int = 8) -> int:
    """sc.py:81-146."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    if M.size == 0:
        return 0
    return int(lib().orc_find_plateau_end(_p(M), i64(M.size), i64(cp_len),
                                          i64(-1 if lookahead is None else int(max(1, lookahead))),
                                          i64(smooth_win), None))


def find_minn_peak(M, smooth_win: int = 8, gate_threshold: float = 0.5, search_bounds=None):
    """minn.py:131-205."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    if M.size == 0:
        raise ValueError("Minn metric is empty")
    gate = np.zeros(M.size, dtype=np.uint8); Ms = np.zeros(M.size)
    hb = search_bounds is not None
    lo, hi = (search_bounds if hb else (0, 0))
    pk = lib().orc_find_minn_peak(_p(M), i64(M.size), i64(smooth_win), C.c_double(gate_threshold),
                                  C.c_int(int(hb)), i64(int(lo)), i64(int(hi)), _p(gate), _p(Ms))
    if pk == -2:
        raise ValueError("Minn metric did not produce a positive peak")
    return int(pk), gate.astype(bool), Ms


def find_minn_peak_gated(M, smooth_win: int = 8, gate_mask=None, search_bounds=None) -> int:
    """combined_sc_min.py:212-259."""
    M = np.ascontiguousarray(M, dtype=np.float64)
    if M.size == 0:
        return 0
    if gate_mask is None:
        raise ValueError("Minn peak detection requires S&C gate mask")
    g = np.ascontiguousarray(gate_mask, dtype=np.uint8)
    if g.shape[0] != M.shape[0]:
        raise ValueError("gate_mask must match metric length")
    Ms = np.zeros(M.size)
    hb = search_bounds is not None
    lo, hi = (search_bounds if hb else (0, 0))
    pk = lib().orc_find_minn_peak_gated(_p(M), i64(M.size), i64(smooth_win), _p(g),
                                        C.c_int(int(hb)), i64(int(lo)), i64(int(hi)), _p(Ms))
    if pk == -3:
        raise ValueError("Minn peak detector received empty gate region")
    return int(pk)


def sc_gate(M_sc, threshold: float = 0.6) -> np.ndarray:
    """combined_sc_min.py:337-351."""
    M_sc = np.ascontiguousarray(M_sc, dtype=np.float64)
    g = np.zeros(M_sc.size, dtype=np.uint8)
    lib().orc_sc_gate(_p(M_sc), i64(M_sc.size), C.c_double(threshold), _p(g))
    return g.astype(bool)


# ----------------------------------------------------------------------------- ZC
def matched_filter_correlation(rx, ref):
    """zc_v2.py:244-254 (also zc.py:116-117): returns (corr, sliding energy)."""
    x = np.ascontiguousarray(rx, dtype=np.complex128).reshape(-1)
    r = np.ascontiguousarray(ref, dtype=np.complex128).reshape(-1)
    n = x.size + r.size - 1
    corr = np.zeros(2 * n); en = np.zeros(n)
    lib().orc_zc_matched_filter(_p(x), i64(x.size), _p(r), i64(r.size), _p(corr), _p(en))
    return _cplx(corr), en


def zc_correlation(rx, ref):
    """zc.py:105-130: branch-summed numerator / energy, normalised AFTER the sum."""
    x = _c128_2d(rx)
    num = None; pw = None
    for b in x:
        c, e = matched_filter_correlation(b, ref)
        num = c if num is None else num + c
        pw = e if pw is None else pw + e
    ref_norm = np.sqrt(np.sum(np.abs(ref) ** 2))
    corr = num / (ref_norm * np.sqrt(np.maximum(pw, 0.0) + 1e-12))
    mag = np.abs(corr)
    peak = int(np.argmax(mag))
    return corr, peak, max(peak - len(ref) + 1, 0)


def zc_v2_corr_mag(rx, ref, normalize: bool = True):
    """zc_v2.py:486-498: per-branch normalise, then sum, then magnitude."""
    x = _c128_2d(rx)
    tot = None
    ref_norm = np.sqrt(np.sum(np.abs(ref) ** 2))
    for b in x:
        c, e = matched_filter_correlation(b, ref)
        if normalize:
            c = c / (ref_norm * np.sqrt(np.maximum(e, 1e-12)))
        tot = c if tot is None else tot + c
    return np.abs(tot)


@dataclass
class ZCState:
    corr_mag: np.ndarray
    local_sum: np.ndarray
    corr_scaled: np.ndarray
    thresh_scaled: np.ndarray
    above_threshold: np.ndarray
    metric_valid: np.ndarray


def zc_streaming_detection(corr_mag, window_size=2048, thresh_value=64, thresh_frac_bits=15, min_corr_mag=0.3):
    """zc_v2.py:288-336."""
    m = np.ascontiguousarray(corr_mag, dtype=np.float64)
    n = m.size
    ls = np.zeros(n); v = np.zeros(n, np.uint8); a = np.zeros(n, np.uint8)
    lib().orc_zc_streaming_detection(_p(m), i64(n), i64(window_size), i64(thresh_value), i64(thresh_frac_bits),
                                     C.c_double(min_corr_mag), _p(ls), _p(v), _p(a))
    return ZCState(m, ls, m * float(1 << thresh_frac_bits), ls * float(thresh_value), a.astype(bool), v.astype(bool))


def detect_zc_peaks(state: ZCState, reference_length: int, hysteresis: int = 256, max_ev: int = 256):
    """zc_v2.py:360-450 -> (events int64[n,4]=(peak,gate_start,gate_end,detected_start), values, gate_mask)."""
    n = state.corr_mag.size
    gm = np.zeros(n, np.uint8); ev = np.zeros((max_ev, 4), np.int64); vals = np.zeros(max_ev)
    v = np.ascontiguousarray(state.metric_valid, dtype=np.uint8)
    a = np.ascontiguousarray(state.above_threshold, dtype=np.uint8)
    nev = lib().orc_detect_zc_peaks(_p(state.corr_mag), _p(v), _p(a), i64(n), i64(reference_length),
                                    i64(hysteresis), _p(gm), _p(ev), _p(vals), i64(max_ev))
    return ev[:nev].copy(), vals[:nev].copy(), gm.astype(bool)


def compute_frequency_metric(rx, bin_indices, template_bins, template_energy, n_fft=2048, cp=512):
    """zc_freq.py:62-99."""
    x = _c128_2d(rx)
    nb, L = x.shape
    nofs = L - (n_fft + cp) + 1
    if nofs <= 0:
        raise ValueError("Received stream is shorter than a single OFDM symbol.")
    bi = np.ascontiguousarray(bin_indices, dtype=np.int64)
    t = np.ascontiguousarray(template_bins, dtype=np.complex128)
    out = np.zeros(nofs)
    lib().orc_zc_freq_metric(_p(x), i64(nb), i64(L), i64(n_fft), i64(cp), _p(bi), _p(t), i64(bi.size),
                             C.c_double(template_energy), _p(out))
    return out


# ----------------------------------------------------------------------------- sync_aa
def aa_detect_streaming(rx, L=512, threshold=0.15, hysteresis=128, sample_rate=15_360_000.0, max_ev=64):
    """sync_aa.py:421-571 -> dict(P,R,M,valid, ev_i[n,4]=(peak,gate_start,gate_end,frame_start),
    ev_f[n,4]=(P_re,P_im,M_at_peak,cfo_hz))."""
    x = _c128_2d(rx)
    na, n = x.shape
    P = np.zeros(2 * n); R = np.zeros(n); M = np.zeros(n); v = np.zeros(n, np.uint8)
    lib().orc_aa_metric(_p(x), i64(na), i64(n), i64(L), _p(P), _p(R), _p(M), _p(v))
    ev_i = np.zeros((max_ev, 4), np.int64); ev_f = np.zeros((max_ev, 4))
    nev = lib().orc_aa_events(_p(P), _p(M), _p(v), i64(n), i64(L), C.c_double(threshold), i64(hysteresis),
                              C.c_double(sample_rate), _p(ev_i), _p(ev_f), i64(max_ev))
    return dict(P=_cplx(P), R=R, M=M, valid=v.astype(bool), ev_i=ev_i[:nev].copy(), ev_f=ev_f[:nev].copy())


# ----------------------------------------------------------------------------- minn_rtl
def minn_rtl_streaming_metric(rx, *, smooth_shift, threshold_value, threshold_frac_bits, quarter_len=512):
    """minn_rtl.py:667-733 -> dict of the 8 state arrays."""
    x = _c128_2d(rx)
    nb, n = x.shape
    f = lambda: np.zeros(n)
    d = dict(corr_total=f(), corr_positive=f(), smooth_metric=f(), energy_total=f(), corr_scaled=f(),
             energy_scaled=f(), metric_valid=np.zeros(n, np.uint8), above=np.zeros(n, np.uint8))
    lib().orc_minn_rtl_metric(_p(x), i64(nb), i64(n), i64(quarter_len), i64(smooth_shift), i64(threshold_value),
                              i64(threshold_frac_bits), *[_p(d[k]) for k in d])
    d["metric_valid"] = d["metric_valid"].astype(bool); d["above"] = d["above"].astype(bool)
    return d


def detect_minn_rtl(state: dict, *, hysteresis: int, timing_offset: int, max_ev: int = 256):
    """minn_rtl.py:750-825 -> (events int64[n,4]=(peak,detected,seg_lo,seg_hi), gate_segments int64[m,2])."""
    n = state["corr_positive"].size
    ev = np.zeros((max_ev, 4), np.int64); sg = np.zeros((max_ev, 2), np.int64); nseg = i64(0)
    a = np.ascontiguousarray(state["above"], dtype=np.uint8)
    v = np.ascontiguousarray(state["metric_valid"], dtype=np.uint8)
    cp = state["corr_positive"]
    if cp.dtype == np.int64:
        nev = lib().orc_detect_minn_rtl_int(_p(cp), _p(a), _p(v), i64(n), i64(hysteresis), i64(timing_offset),
                                            _p(ev), _p(sg), i64(max_ev), C.byref(nseg))
    else:
        cp = np.ascontiguousarray(cp, dtype=np.float64)
        nev = lib().orc_detect_minn_rtl(_p(cp), _p(a), _p(v), i64(n), i64(hysteresis), i64(timing_offset),
                                        _p(ev), _p(sg), i64(max_ev), C.byref(nseg))
    return ev[:nev].copy(), sg[:nseg.value].copy()


def minn_rtl_int(iq, *, smooth_shift, threshold_value, threshold_frac_bits, quarter_len=512, lag_extra=0):
    """Integer model of ref/minn_antenna_path.sv + ref/minn_preamble_detector.sv:247-325.
    iq: int16 (antennas, n, 2)."""
    q = np.ascontiguousarray(iq, dtype=np.int16)
    if q.ndim == 2:
        q = q[np.newaxis]
    nb, n, _ = q.shape
    f = lambda: np.zeros(n, np.int64)
    d = dict(corr_total=f(), corr_positive=f(), smooth_metric=f(), energy_total=f(),
             metric_valid=np.zeros(n, np.uint8), above=np.zeros(n, np.uint8))
    lib().orc_minn_rtl_int(_p(q), i64(nb), i64(n), i64(quarter_len), i64(smooth_shift), i64(threshold_value),
                           i64(threshold_frac_bits), i64(lag_extra), *[_p(d[k]) for k in d])
    d["metric_valid"] = d["metric_valid"].astype(bool); d["above"] = d["above"].astype(bool)
    return d


# ----------------------------------------------------------------------------- scale path
def metric_prefix_c64(x_c64: np.ndarray, n_fft: int, kind: int, want_pr: bool = False):
    """Closed forms of SURVEY.md Appendix B in float64 on a complex64 row (kind 0 S&C, 1 S&C both
    halves, 2 Minn).  Used for BASELINE-size checks and as bench.py's cpu_baseline 'port'."""
    x = np.ascontiguousarray(x_c64, dtype=np.complex64).reshape(-1)
    n = max(x.size - n_fft + 1, 0)
    M = np.zeros(n)
    P = np.zeros(2 * n) if want_pr else None
    R = np.zeros(n) if want_pr else None
    lib().orc_metric_prefix_c64(_p(x), i64(x.size), i64(n_fft), C.c_int(kind), _p(M),
                                _p(P) if want_pr else None, _p(R) if want_pr else None)
    return (M, _cplx(P), R) if want_pr else M


# ----------------------------------------------------------------------------- channel / impairments / CP-CFO
# numpy restatements (small inputs only): these stages are float64 library calls in the reference.
def channel_apply(tx, cir, snr_db: float, unit_noise):
    """channel.apply_channel (channel.py:78-98) with the AWGN of channel._compute_awgn_noise (channel.py:51-75) expressed on
    given unit-normal draws: noise = noise_std * unit_noise, noise_std = sqrt(mean|faded|^2 / snr / 2) per branch row.
    cir: (branches, taps) or None -> (branches, L)."""
    tx = np.asarray(tx)
    faded = tx[np.newaxis, :] if cir is None else np.stack([np.convolve(tx, taps, mode="full") for taps in np.atleast_2d(cir)])
    p = np.mean(np.abs(faded) ** 2, axis=1, keepdims=True)
    std = np.sqrt(p / (10 ** (snr_db / 10)) / 2)
    noise = std * np.asarray(unit_noise).reshape(faded.shape)
    noise[p.squeeze(axis=1) == 0] = 0
    return faded + noise


def unit_noise_like_reference(seed: int, shape) -> np.ndarray:
    """The draws channel._compute_awgn_noise makes from np.random.default_rng(seed): real part first, then imaginary."""
    rng = np.random.default_rng(seed)
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def apply_cfo(x, cfo_hz: float, fs_hz: float):
    """core.apply_cfo (core.py:123-138)."""
    x = np.asarray(x)
    n = np.arange(x.shape[-1], dtype=float)
    return x * np.exp(1j * 2 * np.pi * cfo_hz * n / fs_hz)


def quantize_adc(x, full_scale: float, bits: int = 12):
    """sync_aa.quantize_adc (sync_aa.py:263-291) -> (quantised complex, integer codes [.., 2])."""
    levels = 2 ** (bits - 1)
    def q(v):
        return np.round(np.clip(v / full_scale, -1.0, 1.0 - 1.0 / levels) * levels)
    qr, qi = q(np.real(x)), q(np.imag(x))
    return (qr + 1j * qi) / levels * full_scale, np.stack((qr, qi), axis=-1).astype(np.int16)


def aa_single_test(tx, cir, snr_db: float, full_scale_ratio: float, half_len: int, unit_noise, cfo_hz: float = 500.0,
                   fs_hz: float = 15_360_000.0, true_start: int = 500):
    """One case of the sync_aa grid after the transmit side -- sync_aa.run_single_test, sync_aa.py:716-823: per-antenna channel +
    AWGN (:577-634, on given unit-normal draws), CFO (:637-645), full scale = rms x ratio (:727-728), clipping statistics
    (:294-315), 12-bit quantiser (:263-291), detector (:421-571), strongest event (:741) -> dict of the TestResult fields.
    cir: (antennas, taps); unit_noise: (antennas, n_out) complex."""
    rx = channel_apply(tx, cir, snr_db, unit_noise)
    rx = apply_cfo(rx, cfo_hz, fs_hz)
    rms = np.sqrt(np.mean(np.abs(rx) ** 2))
    full_scale = rms * full_scale_ratio
    flat = rx.reshape(-1)
    clip = np.sum((np.abs(flat.real) >= full_scale) | (np.abs(flat.imag) >= full_scale)) / flat.size
    eff = max(0.0, 12 + np.log2(rms / full_scale)) if full_scale > 0 else 0.0
    q = np.stack([quantize_adc(r, full_scale, 12)[0] for r in rx])
    d = aa_detect_streaming(q, half_len, 0.15, 128, fs_hz)
    out = dict(clipping_pct=100.0 * clip, effective_bits=eff)
    if len(d["ev_i"]):
        k = int(np.argmax(d["ev_f"][:, 2]))                  # max M_at_peak, first on ties
        out.update(detected=True, num_events=len(d["ev_i"]), timing_error=int(d["ev_i"][k, 3]) - true_start,
                   cfo_estimated_hz=float(d["ev_f"][k, 3]), metric_peak=float(d["ev_f"][k, 2]))
    else:
        out.update(detected=False, num_events=0, timing_error=0, cfo_estimated_hz=0.0,
                   metric_peak=float(np.max(d["M"])) if np.any(d["valid"]) else 0.0)
    return out


def peak_noise_statistics(metric, peak_idx: int, pre_pad: int = 1337, guard: int = 500):
    """Peak value, peak / mean(noise), peak / max(noise) with the noise floor taken outside [peak - guard, peak + guard) and past
    the leading pad -- minn.py:841-858 == minn_rtl.py:1561-1580."""
    m = np.asarray(metric, dtype=np.float64)
    keep = np.ones(m.size, dtype=bool)
    keep[max(0, peak_idx - guard):min(m.size, peak_idx + guard)] = False
    keep[:pre_pad] = False
    noise = m[keep]
    peak = float(m[peak_idx])
    if noise.size == 0:
        return peak, float("inf"), float("inf")
    avg, mx = float(np.mean(noise)), float(np.max(noise))
    return peak, (peak / avg if avg > 0 else float("inf")), (peak / mx if mx > 0 else float("inf"))


def block_length_point(tx, frame_len: int, cir, snr_db: float, symbol_len: int, unit_noise, cfo_hz: float = 1000.0,
                       fs_hz: float = 30.72e6, pre_pad: int = 1337, delay: int = 0):
    """One point of minn.compare_block_lengths after the transmit side (minn.py:812-858): channel + CFO, parameterised Minn
    metric, arg-max over the first frame region, noise statistics -> (peak, par, pmr, timing_error)."""
    rx = apply_cfo(channel_apply(tx, cir, snr_db, unit_noise), cfo_hz, fs_hz)
    M, _, _ = minn_streaming_metric(rx, symbol_len)
    end = pre_pad + frame_len + frame_len // 2
    pk = int(np.argmax(M[:min(end, M.size)]))
    peak, par, pmr = peak_noise_statistics(M, pk, pre_pad)
    return peak, par, pmr, pk - (pre_pad + delay + symbol_len // 4)


def q_value_point(tx, cir, snr_db: float, quarter_len: int, unit_noise, cfo_hz: float = 1000.0, fs_hz: float = 30.72e6,
                  pre_pad: int = 1337, delay: int = 0, cp_len: int = 512, smooth_shift: int = 3, threshold_value: int = 3276,
                  frac_bits: int = 15, hysteresis: int = 2, timing_offset: int = 0):
    """One point of minn_rtl.compare_q_values after the transmit side (minn_rtl.py:1535-1580): channel + CFO, the RTL-style metric
    and detector at segment length Q, first event (else arg-max of the smoothed metric), noise statistics of corr_positive
    -> (peak, par, pmr, timing_error)."""
    rx = apply_cfo(channel_apply(tx, cir, snr_db, unit_noise), cfo_hz, fs_hz)
    st = minn_rtl_streaming_metric(rx, smooth_shift=smooth_shift, threshold_value=threshold_value, threshold_frac_bits=frac_bits,
                                   quarter_len=quarter_len)
    ev, _ = detect_minn_rtl(st, hysteresis=hysteresis, timing_offset=timing_offset)
    target = pre_pad + delay + 5 * quarter_len + cp_len
    if len(ev):
        pk, terr = int(ev[0, 0]), int(ev[0, 1]) - target
    else:
        pk = int(np.argmax(st["smooth_metric"]))
        terr = pk - target
    peak, par, pmr = peak_noise_statistics(st["corr_positive"], pk, pre_pad)
    return peak, par, pmr, terr


def wire_pack_hex24(iq):
    """docs/preamble_test_vector.hex: one 24-bit word per sample, {Re[11:0], Im[11:0]}, Re in the upper 12 bits."""
    q = np.asarray(iq, dtype=np.int64)
    return ((q[..., 0] & 0xFFF) << 12) | (q[..., 1] & 0xFFF)


def wire_pack_axis48(iq):
    """ref/test_minn_preamble_detector.py:41-47 (_pack_axis_samples): {ch1_q, ch1_i, ch0_q, ch0_i} x 12 bits, ch0_i lowest.
    iq: int [2, n, 2]."""
    q = np.asarray(iq, dtype=np.int64) & 0xFFF
    return q[0, :, 0] | (q[0, :, 1] << 12) | (q[1, :, 0] << 24) | (q[1, :, 1] << 36)


def _sext12(v):
    v = np.asarray(v, dtype=np.int64) & 0xFFF
    return np.where(v >= 2048, v - 4096, v).astype(np.int16)


def wire_unpack_hex24(words):
    w = np.asarray(words, dtype=np.int64)
    return np.stack((_sext12(w >> 12), _sext12(w)), axis=-1)


def wire_unpack_axis48(words):
    w = np.asarray(words, dtype=np.int64)
    return np.stack((np.stack((_sext12(w), _sext12(w >> 12)), axis=-1), np.stack((_sext12(w >> 24), _sext12(w >> 36)), axis=-1)))


def _cp_P(x, d, n_fft, w):
    return np.sum(x[:, d:d + w] * np.conj(x[:, d + n_fft:d + n_fft + w]))


def estimate_cfo_from_cp(rx, start, n_fft, cp_len, fs_hz):
    """core.py:179-196."""
    x = np.atleast_2d(np.asarray(rx))
    return float(-np.angle(_cp_P(x, start, n_fft, cp_len)) * fs_hz / (2 * np.pi * n_fft))


def estimate_cfo_from_cp_robust(rx, start, n_fft, cp_len, fs_hz, span=None, win_len=None):
    """core.py:199-230."""
    x = np.atleast_2d(np.asarray(rx))
    L = x.shape[1]
    span = cp_len // 2 if span is None else int(max(0, span))
    win = cp_len // 2 if win_len is None else int(max(1, win_len))
    d_lo, d_hi = max(0, start - span), min(L - (n_fft + win), start + span)
    if d_hi <= d_lo:
        return estimate_cfo_from_cp(x, start, n_fft, min(cp_len, win), fs_hz)
    P = 0j
    for d in range(d_lo, d_hi):
        P += _cp_P(x, d, n_fft, win)
    return float(-np.angle(P) * fs_hz / (2 * np.pi * n_fft))


def estimate_cfo_from_cp_peak_with_index(rx, start, n_fft, cp_len, fs_hz, span=None):
    """core.py:233-308 (first maximum of |P(d)|, strict >) -> (cfo_hz, best_d)."""
    x = np.atleast_2d(np.asarray(rx))
    L = x.shape[1]
    span = cp_len // 2 if span is None else int(max(0, span))
    d_lo, d_hi = max(0, start - span), min(L - (n_fft + cp_len), start + span)
    if d_hi <= d_lo:
        return estimate_cfo_from_cp(x, start, n_fft, cp_len, fs_hz), start
    best, bm, bd = 0j, -1.0, d_lo
    for d in range(d_lo, d_hi):
        P = _cp_P(x, d, n_fft, cp_len)
        if abs(P) > bm:
            bm, best, bd = abs(P), P, d
    return float(-np.angle(best) * fs_hz / (2 * np.pi * n_fft)), int(bd)


def rx_chain(rx, pilot_cp_start: int, cfo_hz: float, pilot_used, data_used, fs=30.72e6, n_fft=2048, cp_len=512):
    """sc.py:286-309 / minn.py:546-568 with core.apply_cfo, ofdm_fft_used (core.py:171-176), ls_channel_estimate, equalize
    (core.py:339-345), align_complex_gain (core.py:357-362), evm_rms_db (core.py:365-370) and
    estimate_timing_offset_from_phase_slope (core.py:443-469)."""
    x = apply_cfo(np.atleast_2d(np.asarray(rx)), -cfo_hz, fs)
    eff = np.mean(x, axis=0)
    nu = len(pilot_used); half = nu // 2
    idx = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
    def used(td):
        return np.fft.fftshift(np.fft.fft(td, n=n_fft))[(n_fft // 2 + idx) % n_fft]
    yp = used(eff[pilot_cp_start + cp_len: pilot_cp_start + cp_len + n_fft])
    h = yp / (np.asarray(pilot_used) + 1e-9)
    k = idx.astype(np.float64)
    phi = np.unwrap(np.angle(h))
    kz, pz = k - np.mean(k), phi - np.mean(phi)
    slope = float(np.sum(kz * pz) / (float(np.sum(kz * kz)) + 1e-12))
    d0 = pilot_cp_start + cp_len + n_fft
    yd = used(eff[d0 + cp_len: d0 + cp_len + n_fft])
    xh = yd / (h + 1e-9)
    ref = np.asarray(data_used)
    g = np.vdot(xh, ref) / (np.vdot(xh, xh) + 1e-12)
    xa = xh * g
    evm = float(np.sqrt(np.mean(np.abs(xa - ref) ** 2) / np.mean(np.abs(ref) ** 2)))
    return dict(h_est=h, xhat=xa, gain=g, evm_rms=evm, evm_db=float(20 * np.log10(evm + 1e-12)), slope=slope,
                sto=float(-slope * n_fft / (2 * np.pi)))
