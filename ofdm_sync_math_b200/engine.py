"""Batched device API over libofdmsync (torch tensors in, torch tensors out).

torch is plumbing only: device memory, streams, (in dist.py) process groups.  Every array
operation on the hot path is a hand-written sm_100a kernel behind the C ABI of include/ofdmsync.h.
Inputs: numpy or torch; shapes (L,), (branches, L) or (frames, branches, L); complex64 / complex128,
or int16 IQ with a trailing axis of 2.  Branches are summed before the metric, frames are independent.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L

KINDS = {"sc": L.OFS_SC, "sc_both": L.OFS_SC_BOTH, "minn": L.OFS_MINN, "aa": L.OFS_AA}
PATHS = {"auto": L.OFS_PATH_AUTO, "stripe": L.OFS_PATH_STRIPE, "tile": L.OFS_PATH_TILE, "array": L.OFS_PATH_ARRAY}


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise L.OfsError("ofdm_sync_math_b200 needs a CUDA device (no CPU fallback exists)")
    return torch.device("cuda", torch.cuda.current_device())


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: torch.Tensor | None) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def to_device(rx) -> tuple[torch.Tensor, int, bool]:
    """-> (tensor [F, B, L(,2)] contiguous on the GPU, ofs dtype code, input_was_numpy)."""
    was_numpy = not isinstance(rx, torch.Tensor)
    t = torch.as_tensor(np.ascontiguousarray(rx)) if was_numpy else rx
    if t.dtype == torch.int16:
        if t.shape[-1] != 2:
            raise ValueError("int16 input must have a trailing I/Q axis of length 2")
        code, nd = L.OFS_IQ16, t.dim() - 1
    elif t.dtype == torch.complex64:
        code, nd = L.OFS_C64, t.dim()
    elif t.dtype == torch.complex128:
        code, nd = L.OFS_C128, t.dim()
    elif t.dtype in (torch.float32, torch.float64, torch.int32, torch.int64):
        t = t.to(torch.complex128)
        code, nd = L.OFS_C128, t.dim()
    else:
        raise TypeError(f"unsupported sample dtype {t.dtype}")
    if nd == 1:
        t = t[None, None]
    elif nd == 2:
        t = t[None]
    elif nd != 3:
        raise ValueError("samples must be 1-D, 2-D (branches, L) or 3-D (frames, branches, L)")
    t = t.to(_device(), non_blocking=True).contiguous()
    return t, code, was_numpy


@dataclass
class MetricOut:
    M: torch.Tensor          # [F, out_len]
    P: torch.Tensor | None   # [F, out_len] complex
    R: torch.Tensor | None
    chunk_max: torch.Tensor | None = None
    path: str = "tile"


def metric_out_len(kind: str, symbol_len: int, n: int) -> int:
    return n if kind == "aa" else max(n - symbol_len + 1, 0)


def metric(rx, kind: str, symbol_len: int, *, want_pr: bool = True, out_f64: bool | None = None,
           path: str = "auto", store_mode: int = 0, want_chunk_max: bool = False, tma_mode: int = 2) -> MetricOut:
    """Timing metric M (+P, R) of sc.py:42-78 / combined_sc_min.py:116-164 / minn.py:59-112,697-751 /
    sync_aa.py:458-493 for a batch of frames."""
    x, code, _ = to_device(rx)
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    dev = x.device
    if out_f64 is None:
        out_f64 = code == L.OFS_C128
    out_len = metric_out_len(kind, symbol_len, n)
    rdt = torch.float64 if out_f64 else torch.float32
    cdt = torch.complex128 if out_f64 else torch.complex64
    if out_len == 0 or F == 0:
        e = torch.zeros((F, 0), dtype=rdt, device=dev)
        return MetricOut(e, torch.zeros((F, 0), dtype=cdt, device=dev) if want_pr else None, e.clone() if want_pr else None)
    d = L.MetricDesc(kind=KINDS[kind], in_dtype=code, out_f64=int(out_f64), path=PATHS[path], symbol_len=int(symbol_len),
                     n_branches=B, n_frames=F, n_samples=n, x_frame_stride=B * n, x_branch_stride=n,
                     out_stride=0, store_mode=store_mode, reserved=1 if tma_mode == 1 else 0)
    lib = L.lib()
    stripe_ok = not out_f64 and bool(lib.ofs_metric_stripe_ok(C.byref(d), _ptr(x), None)) and not (B > 1 and want_pr)
    # auto keeps P / R requests on the precise kernels; path="stripe" serves them from the fast kernel (float32 windows)
    use_stripe = stripe_ok and (path == "stripe" or (path == "auto" and not want_pr))
    if path == "stripe" and not use_stripe:
        raise L.OfsError("stripe path cannot serve this request (needs c64/iq16, lag in {256,512,1024}, float32 outputs; several "
                         "branches: kind sc / sc_both / minn, M only, frame and branch pitch multiples of 128 bytes)")
    if use_stripe:
        # causal-time rows: element d of a frame lives at column toff + d, so the kernel's 16-byte
        # vector / bulk stores (which work in t = d + toff) are aligned.
        toff = 0 if kind == "aa" else symbol_len - 1
        pitch = (n + 3) // 4 * 4
        buf = torch.empty((F, pitch), dtype=torch.float32, device=dev)
        M = buf[:, toff:toff + out_len]
        Pb = torch.empty((F, pitch), dtype=torch.complex64, device=dev) if want_pr else None
        Rb = torch.empty((F, pitch), dtype=torch.float32, device=dev) if want_pr else None
        P = None if Pb is None else Pb[:, toff:toff + out_len]
        R = None if Rb is None else Rb[:, toff:toff + out_len]
        cm = None
        cm_stride = 0
        if want_chunk_max:
            cm_stride = (n + 255) // 256
            cm = torch.zeros((F, cm_stride), dtype=torch.float32, device=dev)
        d.out_stride = pitch
        d.path = L.OFS_PATH_STRIPE
        L.check(lib.ofs_metric(C.byref(d), _ptr(x), C.c_void_p(M.data_ptr()), C.c_void_p(0 if P is None else P.data_ptr()),
                               C.c_void_p(0 if R is None else R.data_ptr()), _ptr(cm), C.c_int64(cm_stride), _stream()), "ofs_metric(stripe)")
        return MetricOut(M, P, R, cm, "stripe")
    array_ok = kind == "aa" and not out_f64 and bool(lib.ofs_metric_array_ok(C.byref(d), _ptr(x)))
    if path == "array" and not array_ok:
        raise L.OfsError("array path cannot serve this request (kind aa, c64/iq16, float32 out, L in {128,256,512,1024}, 16-byte rows)")
    if path == "array" or (path == "auto" and array_ok and B >= 2):
        # antenna-array kernel (metric_array.cu): rows padded to an even pitch for its vector stores
        pitch = (out_len + 1) // 2 * 2
        Mb = torch.empty((F, pitch), dtype=torch.float32, device=dev)
        Pb = torch.empty((F, pitch), dtype=torch.complex64, device=dev) if want_pr else None
        Rb = torch.empty((F, pitch), dtype=torch.float32, device=dev) if want_pr else None
        d.out_stride = pitch
        d.path = L.OFS_PATH_ARRAY if path == "array" else L.OFS_PATH_AUTO
        L.check(lib.ofs_metric(C.byref(d), _ptr(x), _ptr(Mb), _ptr(Pb), _ptr(Rb), None, C.c_int64(0), _stream()), "ofs_metric(array)")
        cut = lambda t: None if t is None else t[:, :out_len]
        return MetricOut(cut(Mb), cut(Pb), cut(Rb), None, "array")
    M = torch.empty((F, out_len), dtype=rdt, device=dev)
    P = torch.empty((F, out_len), dtype=cdt, device=dev) if want_pr else None
    R = torch.empty((F, out_len), dtype=rdt, device=dev) if want_pr else None
    d.out_stride = out_len
    d.path = L.OFS_PATH_TILE
    L.check(lib.ofs_metric(C.byref(d), _ptr(x), _ptr(M), _ptr(P), _ptr(R), None, C.c_int64(0), _stream()), "ofs_metric(tile)")
    return MetricOut(M, P, R, None, "tile")


# ------------------------------------------------------------------------------------------- detectors
def _rows(M: torch.Tensor) -> tuple[L.Rows, torch.Tensor]:
    if M.dim() == 1:
        M = M[None]
    if M.dtype not in (torch.float32, torch.float64):
        M = M.to(torch.float64)
    if M.stride(-1) != 1:
        M = M.contiguous()
    if not M.is_cuda:
        M = M.to(_device())
    return L.Rows(C.c_void_p(M.data_ptr()), int(M.dtype == torch.float64), 0, M.shape[0], M.shape[1], M.stride(0) if M.shape[0] > 1 else max(M.shape[1], M.stride(0))), M


def find_plateau_end(M: torch.Tensor, cp_len: int, lookahead: int | None = None, smooth_win: int = 8,
                     chunk_max: torch.Tensor | None = None, toff: int = 0) -> torch.Tensor:
    """sc.py:81-146 for every row -> int64[rows].  chunk_max/toff (from the stripe metric) prune the argmax pass."""
    rows, M = _rows(M)
    out = torch.zeros(M.shape[0], dtype=torch.int64, device=M.device)
    L.check(L.lib().ofs_find_plateau_end_pruned(C.byref(rows), _ptr(chunk_max), C.c_int64(0 if chunk_max is None else chunk_max.stride(0)),
                                                int(toff), int(cp_len), -1 if lookahead is None else int(max(1, lookahead)),
                                                int(smooth_win), _ptr(out), _stream()), "ofs_find_plateau_end")
    return out


def find_minn_peak(M: torch.Tensor, smooth_win: int = 8, gate_threshold: float = 0.5, search_bounds=None,
                   want_ms: bool = False, chunk_max: torch.Tensor | None = None, toff: int = 0):
    """minn.py:131-205 for every row -> (peak int64[rows], gate_span int64[rows,2], Ms or None)."""
    rows, M = _rows(M)
    if want_ms and not M.is_contiguous():
        rows, M = _rows(M.contiguous())
    peak = torch.zeros(M.shape[0], dtype=torch.int64, device=M.device)
    span = torch.zeros((M.shape[0], 2), dtype=torch.int64, device=M.device)
    Ms = torch.empty(M.shape, dtype=M.dtype, device=M.device) if want_ms else None
    hb = search_bounds is not None
    lo, hi = (int(search_bounds[0]), int(search_bounds[1])) if hb else (0, 0)
    L.check(L.lib().ofs_find_minn_peak_pruned(C.byref(rows), _ptr(chunk_max), C.c_int64(0 if chunk_max is None else chunk_max.stride(0)),
                                              int(toff), int(smooth_win), C.c_double(gate_threshold), int(hb), C.c_int64(lo),
                                              C.c_int64(hi), _ptr(peak), _ptr(span), _ptr(Ms), _stream()), "ofs_find_minn_peak")
    return peak, span, Ms


def sc_gate(M_sc: torch.Tensor, threshold: float = 0.6, chunk_max: torch.Tensor | None = None, toff: int = 0) -> torch.Tensor:
    """combined_sc_min.py:337-351 -> uint8 gate [rows, n].  chunk_max / toff (from the stripe metric) skip the chunks that
    cannot reach the gate level."""
    rows, M_sc = _rows(M_sc)
    gate = torch.empty(M_sc.shape, dtype=torch.uint8, device=M_sc.device)
    L.check(L.lib().ofs_sc_gate_pruned(C.byref(rows), _ptr(chunk_max), C.c_int64(0 if chunk_max is None else chunk_max.stride(0)), int(toff),
                                       C.c_double(threshold), _ptr(gate), C.c_int64(gate.stride(0) if gate.shape[0] > 1 else gate.shape[1]),
                                       _stream()), "ofs_sc_gate")
    return gate


def find_minn_peak_gated(M: torch.Tensor, smooth_win: int, gate: torch.Tensor, search_bounds=None) -> torch.Tensor:
    """combined_sc_min.py:183-259 -> int64[rows] (-1 empty metric, -3 empty gate)."""
    rows, M = _rows(M)
    g = gate if isinstance(gate, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(gate))
    if g.dim() == 1:
        g = g[None]
    g = g.to(device=M.device, dtype=torch.uint8).contiguous()
    peak = torch.zeros(M.shape[0], dtype=torch.int64, device=M.device)
    hb = search_bounds is not None
    lo, hi = (int(search_bounds[0]), int(search_bounds[1])) if hb else (0, 0)
    L.check(L.lib().ofs_find_minn_peak_gated(C.byref(rows), int(smooth_win), _ptr(g), C.c_int64(g.shape[1]), int(hb),
                                             C.c_int64(lo), C.c_int64(hi), _ptr(peak), _stream()), "ofs_find_minn_peak_gated")
    return peak


def combined_peak(M_minn: torch.Tensor, M_sc: torch.Tensor, chunk_max_sc: torch.Tensor, toff: int, threshold: float = 0.6,
                  smooth_win: int = 16, search_bounds=None):
    """S&C-gated Minn peak (combined_sc_min.py:183-259, 337-351) in one launch, no gate array (ofs_combined_peak).
    float32 metrics of equal shape from the stripe kernel + the S&C chunk maxima -> (peak int64[rows], gate_span int64[rows, 2])."""
    rm, M_minn = _rows(M_minn)
    rs, M_sc = _rows(M_sc)
    peak = torch.zeros(M_sc.shape[0], dtype=torch.int64, device=M_sc.device)
    span = torch.zeros((M_sc.shape[0], 2), dtype=torch.int64, device=M_sc.device)
    hb = search_bounds is not None
    lo, hi = (int(search_bounds[0]), int(search_bounds[1])) if hb else (0, 0)
    L.check(L.lib().ofs_combined_peak(C.byref(rm), C.byref(rs), _ptr(chunk_max_sc), C.c_int64(chunk_max_sc.stride(0)), int(toff),
                                      C.c_double(threshold), int(smooth_win), int(hb), C.c_int64(lo), C.c_int64(hi), _ptr(peak),
                                      _ptr(span), _stream()), "ofs_combined_peak")
    return peak, span


def argmax(M: torch.Tensor) -> torch.Tensor:
    rows, M = _rows(M)
    out = torch.zeros(M.shape[0], dtype=torch.int64, device=M.device)
    L.check(L.lib().ofs_argmax(C.byref(rows), _ptr(out), _stream()), "ofs_argmax")
    return out


_EVENT_NP = np.dtype([("peak_index", "<i8"), ("gate_start", "<i8"), ("gate_end", "<i8"), ("aux", "<i8"), ("value", "<f8"),
                      ("p_re", "<f8"), ("p_im", "<f8"), ("cfo", "<f8"), ("closed", "<i4"), ("reserved", "<i4")])
assert _EVENT_NP.itemsize == C.sizeof(L.Event)


class EventOverflow(L.OfsError):
    """A row produced more gates than the slots of the event buffer it was given (the reference returns unbounded lists)."""


def _events_to_numpy(ev: torch.Tensor, cnt: torch.Tensor, cap: int = L.OFS_MAX_EVENTS, overflow: str = "raise") -> list[np.ndarray]:
    """Device event buffers -> one structured array per row.  Only the occupied part of the buffers crosses PCIe (a row holds
    up to `cap` events of 72 bytes; typical captures use a handful).  n_events carries the TRUE number of gates of a row: when
    it exceeds the slots, overflow="raise" raises EventOverflow (with .needed = the largest count) and overflow="truncate" keeps
    the first `cap` events -- nothing is ever cut silently."""
    craw = cnt.cpu().numpy()
    if overflow == "raise" and craw.size and int(craw.max()) > cap:
        r = int(np.argmax(craw))
        err = EventOverflow(f"row {r} has {int(craw[r])} gates but the event buffer holds {cap}")
        err.needed = int(craw.max())
        raise err
    c = np.minimum(craw, cap)
    rows = ev.shape[0]
    mc = int(c.max()) if rows else 0
    if mc == 0:
        return [np.zeros(0, dtype=_EVENT_NP) for _ in range(rows)]
    esz = C.sizeof(L.Event)
    raw = ev.view(rows, cap, esz)[:, :mc].contiguous().cpu().numpy().view(_EVENT_NP).reshape(rows, mc)
    return [raw[i, : int(c[i])] for i in range(rows)]


def _event_buffers(n_rows: int, dev, want_flat: bool = False, cap: int = L.OFS_MAX_EVENTS):
    """Event slots [n_rows, cap * 72 B] and counts int32[n_rows] as two views of ONE allocation, so that a multi-GPU
    caller ships both with a single all_gather (flat uint8 tensor, want_flat=True)."""
    row_b = cap * C.sizeof(L.Event)
    flat = torch.zeros(n_rows * (row_b + 4), dtype=torch.uint8, device=dev)
    ev = flat[: n_rows * row_b].view(n_rows, row_b)
    cnt = flat[n_rows * row_b:].view(torch.int32)
    return (ev, cnt, flat) if want_flat else (ev, cnt)


def _with_event_buffers(n_rows: int, dev, launch, max_events: int | None):
    """Run `launch(ev, cnt, cap)` and decode its events.  max_events=None (the drop-ins): start with OFS_MAX_EVENTS slots and,
    when a row holds more gates than that, repeat the launch with exactly as many slots as the fullest row needs -- lists are
    unbounded like the reference's.  An explicit max_events is a hard cap: overflow raises EventOverflow."""
    cap = L.OFS_MAX_EVENTS if max_events is None else int(max_events)
    while True:
        ev, cnt = _event_buffers(n_rows, dev, cap=cap)
        launch(ev, cnt, cap)
        try:
            return _events_to_numpy(ev, cnt, cap)
        except EventOverflow as e:
            if max_events is not None:
                raise
            cap = e.needed


def aa_events(M: torch.Tensor, P: torch.Tensor, half_len: int, threshold: float, hysteresis: int, sample_rate: float,
              max_events: int | None = None):
    """sync_aa.py:495-568 -> list (per row) of structured event arrays (unbounded; max_events: hard cap, EventOverflow beyond)."""
    rows, M = _rows(M)
    if P.dim() == 1:
        P = P[None]
    P = P.to(torch.complex128 if M.dtype == torch.float64 else torch.complex64).contiguous()

    def launch(ev, cnt, cap):
        L.check(L.lib().ofs_aa_events(C.byref(rows), _ptr(P), int(half_len), C.c_double(threshold), int(hysteresis),
                                      C.c_double(sample_rate), _ptr(ev), _ptr(cnt), int(cap), _stream()), "ofs_aa_events")
    return _with_event_buffers(M.shape[0], M.device, launch, max_events)


def zc_streaming_detection(corr_mag: torch.Tensor, window: int, thresh_value: int, frac_bits: int, min_corr_mag: float):
    """zc_v2.py:288-336 -> (local_sum, valid uint8, above uint8)."""
    rows, m = _rows(corr_mag)
    m = m.contiguous()
    rows, m = _rows(m)
    ls = torch.empty_like(m)
    valid = torch.empty(m.shape, dtype=torch.uint8, device=m.device)       # every element is written by the kernel
    above = torch.empty(m.shape, dtype=torch.uint8, device=m.device)
    L.check(L.lib().ofs_zc_streaming_detection(C.byref(rows), int(window), int(thresh_value), int(frac_bits),
                                               C.c_double(min_corr_mag), _ptr(ls), _ptr(valid), _ptr(above),
                                               C.c_int64(m.shape[1]), _stream()), "ofs_zc_streaming_detection")
    return ls, valid, above


def zc_events(corr_mag: torch.Tensor, valid: torch.Tensor, above: torch.Tensor, reference_length: int, hysteresis: int,
              want_gate_mask: bool = True, max_events: int | None = None):
    """zc_v2.py:360-450 -> (events per row, gate_mask uint8 or None)."""
    rows, m = _rows(corr_mag)
    m = m.contiguous()
    rows, m = _rows(m)
    v = valid.to(device=m.device, dtype=torch.uint8).reshape(m.shape).contiguous()
    a = above.to(device=m.device, dtype=torch.uint8).reshape(m.shape).contiguous()
    gm = torch.zeros(m.shape, dtype=torch.uint8, device=m.device) if want_gate_mask else None

    def launch(ev, cnt, cap):
        L.check(L.lib().ofs_zc_events(C.byref(rows), _ptr(v), _ptr(a), C.c_int64(m.shape[1]), int(reference_length), int(hysteresis),
                                      _ptr(ev), _ptr(cnt), int(cap), _ptr(gm), _stream()), "ofs_zc_events")
    return _with_event_buffers(m.shape[0], m.device, launch, max_events), gm


def zc_detect(corr_mag: torch.Tensor, window: int, thresh_value: int, frac_bits: int, min_corr_mag: float, reference_length: int,
              hysteresis: int, max_events: int | None = None):
    """zc_v2 streaming threshold + gate FSM (zc_v2.py:288-336, 360-450) exchanging a bitmask (ofs_zc_detect) -> events per row."""
    rows, m = _rows(corr_mag)
    m = m.contiguous()
    rows, m = _rows(m)
    mstride = (m.shape[1] + 31) // 32
    mask = torch.empty((m.shape[0], mstride), dtype=torch.int32, device=m.device)

    def launch(ev, cnt, cap):
        L.check(L.lib().ofs_zc_detect(C.byref(rows), int(window), int(thresh_value), int(frac_bits), C.c_double(min_corr_mag),
                                      int(reference_length), int(hysteresis), _ptr(mask), C.c_int64(mstride), _ptr(ev), _ptr(cnt),
                                      int(cap), _stream()), "ofs_zc_detect")
    return _with_event_buffers(m.shape[0], m.device, launch, max_events)


def minn_rtl_events(corr_positive: torch.Tensor, valid: torch.Tensor, above: torch.Tensor, hysteresis: int, timing_offset: int,
                    max_events: int | None = None):
    """minn_rtl.py:750-825 -> events per row (closed == 0 marks an unclosed tail segment)."""
    cp = corr_positive if corr_positive.dim() == 2 else corr_positive[None]
    is_int = cp.dtype == torch.int64
    if not is_int:
        cp = cp.to(torch.float64)
    cp = cp.to(_device()).contiguous()
    v = valid.to(device=cp.device, dtype=torch.uint8).reshape(cp.shape).contiguous()
    a = above.to(device=cp.device, dtype=torch.uint8).reshape(cp.shape).contiguous()
    def launch(ev, cnt, cap):
        L.check(L.lib().ofs_minn_rtl_events(_ptr(cp), int(is_int), _ptr(v), _ptr(a), C.c_int64(cp.shape[0]), C.c_int64(cp.shape[1]),
                                            C.c_int64(cp.shape[1]), int(hysteresis), int(timing_offset), _ptr(ev), _ptr(cnt),
                                            int(cap), _stream()), "ofs_minn_rtl_events")
    return _with_event_buffers(cp.shape[0], cp.device, launch, max_events)


# ------------------------------------------------------------------------------------------- sync pipeline
_REC_NP = np.dtype([("timing", "<i8"), ("coarse", "<i8"), ("metric", "<f4"), ("p_re", "<f4"), ("p_im", "<f4"), ("cfo", "<f4"),
                    ("status", "<i4"), ("reserved", "<i4")])
assert _REC_NP.itemsize == C.sizeof(L.SyncRecord)
REC_BYTES = C.sizeof(L.SyncRecord)


def _sync_params(cp_len, smooth_win, sc_delta, gate_threshold, exact, exact_band) -> L.SyncParams:
    return L.SyncParams(cp_len=int(cp_len), smooth_win=int(smooth_win), sc_delta=int(sc_delta), exact=int(bool(exact)),
                        gate_threshold=float(gate_threshold), exact_band=float(exact_band or 0.0))


@dataclass
class SyncOut:
    M: torch.Tensor            # [F, out_len] float32 view
    records: torch.Tensor      # [F, REC_BYTES] uint8 (ofs_sync_record); use records_numpy()
    chunk_max: torch.Tensor

    def records_numpy(self) -> np.ndarray:
        return self.records.cpu().numpy().view(_REC_NP).reshape(-1)


def sync_f64(x: torch.Tensor, kind: str, symbol_len: int, *, cp_len: int = 512, smooth_win: int = 16, sc_delta: int = 16,
             gate_threshold: float = 0.5) -> torch.Tensor:
    """ofs_sync_f64: the sync pipeline entirely in float64 for frames [F, B, L] (int16 IQ: [F, B, L, 2]) -> records [F, REC_BYTES]."""
    if x.dim() != (4 if x.dtype == torch.int16 else 3):
        raise ValueError("sync_f64 takes frames [F, B, L] (int16 IQ: [F, B, L, 2])")
    xd, code, _ = to_device(x)
    F, B, n = xd.shape[0], xd.shape[1], xd.shape[2]
    rec = torch.zeros((F, REC_BYTES), dtype=torch.uint8, device=xd.device)
    d = L.MetricDesc(kind=KINDS[kind], in_dtype=code, out_f64=1, path=L.OFS_PATH_TILE, symbol_len=int(symbol_len), n_branches=B,
                     n_frames=F, n_samples=n, x_frame_stride=B * n, x_branch_stride=n, out_stride=max(n - symbol_len + 1, 0),
                     store_mode=0, reserved=0)
    sp = _sync_params(cp_len, smooth_win, sc_delta, gate_threshold, False, 0.0)
    L.check(L.lib().ofs_sync_f64(C.byref(d), _ptr(xd), C.byref(sp), _ptr(rec), _stream()), "ofs_sync_f64")
    return rec


class SyncPlan:
    """Pre-allocated buffers for repeated ofs_sync calls on device-resident frames [F, L] (or [F, B, L], branches summed)
    complex64 / int16-IQ.  exact=True (default): detector decisions inside the float32 error band of the stripe metric are
    re-evaluated in float64 from the samples, so `timing` equals what the reference finds on its float64 metric
    (records["status"]: OFS_ST_EXACT / _CHANGED / _UNRESOLVED; resolve() settles the rare unresolved frames)."""

    def __init__(self, n_frames: int, n_samples: int, kind: str = "sc", symbol_len: int = 2048, in_dtype: str = "c64",
                 cp_len: int = 512, smooth_win: int = 16, sc_delta: int = 16, gate_threshold: float = 0.5, store_mode: int = 0,
                 tma_mode: int = 2, n_branches: int = 1, exact: bool = True, exact_band: float = 0.0):
        dev = _device()
        self.F, self.n, self.kind, self.N, self.B = n_frames, n_samples, kind, symbol_len, n_branches
        self.code = {"c64": L.OFS_C64, "iq16": L.OFS_IQ16}[in_dtype]
        self.cp_len, self.smooth_win, self.sc_delta, self.gate_threshold = cp_len, smooth_win, sc_delta, gate_threshold
        self.params = _sync_params(cp_len, smooth_win, sc_delta, gate_threshold, exact, exact_band)
        self.toff = symbol_len - 1
        self.pitch = (n_samples + 3) // 4 * 4
        self.out_len = max(n_samples - symbol_len + 1, 0)
        self.buf = torch.empty((n_frames, self.pitch), dtype=torch.float32, device=dev)
        self.M = self.buf[:, self.toff:self.toff + self.out_len]
        self.cm_stride = (n_samples + 255) // 256
        self.cm = torch.zeros((n_frames, self.cm_stride), dtype=torch.float32, device=dev)
        self.rec = torch.zeros((n_frames, REC_BYTES), dtype=torch.uint8, device=dev)
        self.scratch = torch.zeros(4 * n_frames, dtype=torch.int64, device=dev)
        self.desc = L.MetricDesc(kind=KINDS[kind], in_dtype=self.code, out_f64=0, path=L.OFS_PATH_AUTO, symbol_len=symbol_len,
                                 n_branches=n_branches, n_frames=n_frames, n_samples=n_samples, x_frame_stride=n_branches * n_samples,
                                 x_branch_stride=n_samples, out_stride=self.pitch, store_mode=store_mode,
                                 reserved=1 if tma_mode == 1 else 0)

    def run(self, x: torch.Tensor) -> SyncOut:
        assert x.is_cuda and x.is_contiguous() and x.shape[0] == self.F
        L.check(L.lib().ofs_sync(C.byref(self.desc), _ptr(x), C.c_void_p(self.M.data_ptr()), _ptr(self.cm), C.c_int64(self.cm_stride),
                                 C.byref(self.params), _ptr(self.rec), _ptr(self.scratch), _stream()), "ofs_sync")
        return SyncOut(self.M, self.rec, self.cm)

    def records_numpy(self) -> np.ndarray:
        return self.rec.cpu().numpy().view(_REC_NP).reshape(-1)

    def resolve(self, x: torch.Tensor) -> int:
        """Re-run the frames the kernels flagged OFS_ST_UNRESOLVED through the float64 pipeline (ofs_sync_f64) and patch their
        records in place.  Synchronises (reads the status column).  -> number of frames re-run."""
        st = self.rec.view(torch.int32).view(self.F, -1)[:, 8]
        idx = torch.nonzero((st & L.OFS_ST_UNRESOLVED) != 0).flatten()
        if idx.numel() == 0:
            return 0
        xs = x.index_select(0, idx).contiguous()
        if xs.dim() == 2 + (self.code == L.OFS_IQ16):          # [F, L(, 2)] -> [F, 1, L(, 2)]
            xs = xs[:, None]
        rec = sync_f64(xs, self.kind, self.N, cp_len=self.cp_len, smooth_win=self.smooth_win, sc_delta=self.sc_delta,
                       gate_threshold=self.gate_threshold)
        self.rec.index_copy_(0, idx, rec)
        return int(idx.numel())

    def run_detect_only(self, x: torch.Tensor) -> None:
        L.check(L.lib().ofs_sync_detect(C.byref(self.desc), _ptr(x), C.c_void_p(self.M.data_ptr()), _ptr(self.cm),
                                        C.c_int64(self.cm_stride), C.byref(self.params), _ptr(self.rec),
                                        _ptr(self.scratch), _stream()), "ofs_sync_detect")

    def run_metric_only(self, x: torch.Tensor) -> None:
        L.check(L.lib().ofs_metric(C.byref(self.desc), _ptr(x), C.c_void_p(self.M.data_ptr()), None, None, _ptr(self.cm),
                                   C.c_int64(self.cm_stride), _stream()), "ofs_metric")


class HostSync:
    """ofs_sync_host: frames in (pinned) host memory -> M + records in host memory, copies pipelined inside the library."""

    def __init__(self, device: int | None = None):
        self.ctx = C.c_void_p()
        dev = torch.cuda.current_device() if device is None else device
        L.check(L.lib().ofs_ctx_create(C.byref(self.ctx), int(dev)), "ofs_ctx_create")

    def close(self):
        if self.ctx:
            L.lib().ofs_ctx_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, x_host: torch.Tensor, M_host: torch.Tensor | None, records_host: torch.Tensor, *, kind="sc", symbol_len=2048,
            cp_len=512, smooth_win=16, sc_delta=16, gate_threshold=0.5, exact=True, exact_band=0.0):
        F, n = x_host.shape[0], x_host.shape[1]
        if x_host.dtype == torch.int16:
            # int16 IQ [F, n, 2]: ofs_metric_desc strides count SAMPLES (4 bytes each), torch strides count int16 elements
            if x_host.dim() != 3 or x_host.shape[2] != 2 or x_host.stride(2) != 1 or x_host.stride(1) != 2 or x_host.stride(0) % 2:
                raise ValueError("int16 IQ host frames must be [frames, n, 2] with interleaved (I, Q) pairs")
            code, xfs = L.OFS_IQ16, x_host.stride(0) // 2
        elif x_host.dtype == torch.complex64:
            if x_host.dim() != 2 or x_host.stride(1) != 1:
                raise ValueError("complex64 host frames must be [frames, n] with unit sample stride")
            code, xfs = L.OFS_C64, x_host.stride(0)
        else:
            raise TypeError(f"ofs_sync_host takes complex64 or int16 IQ frames, not {x_host.dtype}")
        out_len = max(n - symbol_len + 1, 0)
        if F > 1 and xfs < n:
            raise ValueError("frame stride shorter than a frame")
        if M_host is not None and (M_host.dtype != torch.float32 or M_host.dim() != 2 or M_host.shape[0] < F or M_host.shape[1] < out_len
                                   or M_host.stride(1) != 1 or (F > 1 and M_host.stride(0) < out_len)):
            raise ValueError(f"M_host must be float32 [frames, >= {out_len}] with unit element stride")
        if records_host.numel() * records_host.element_size() < F * REC_BYTES:
            raise ValueError("records_host too small")
        out_stride = M_host.stride(0) if M_host is not None else out_len
        d = L.MetricDesc(kind=KINDS[kind], in_dtype=code, out_f64=0, path=L.OFS_PATH_AUTO, symbol_len=symbol_len, n_branches=1,
                         n_frames=F, n_samples=n, x_frame_stride=xfs, x_branch_stride=n, out_stride=out_stride,
                         store_mode=0, reserved=0)
        sp = _sync_params(cp_len, smooth_win, sc_delta, gate_threshold, exact, exact_band)
        L.check(L.lib().ofs_sync_host(self.ctx, C.byref(d), _ptr(x_host), _ptr(M_host), C.byref(sp), _ptr(records_host)), "ofs_sync_host")
        return records_host.numpy().reshape(-1)[: F * REC_BYTES].view(_REC_NP).reshape(-1)


# ------------------------------------------------------------------------------------------- park
def park_metric(rx, symbol_len: int = 2048, out_f64: bool | None = None):
    """park.py:64-114 -> (M, P, E) tensors [F, n], n = L - 2*(N/2)."""
    x, code, _ = to_device(rx)
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    if out_f64 is None:
        out_f64 = code == L.OFS_C128
    h = symbol_len // 2
    n_out = max(n - 2 * h, 0) if h > 0 else 0
    rdt = torch.float64 if out_f64 else torch.float32
    cdt = torch.complex128 if out_f64 else torch.complex64
    M = torch.empty((F, n_out), dtype=rdt, device=x.device)
    P = torch.empty((F, n_out), dtype=cdt, device=x.device)
    E = torch.empty((F, n_out), dtype=rdt, device=x.device)
    if n_out > 0 and F > 0:
        d = L.MetricDesc(kind=0, in_dtype=code, out_f64=int(out_f64), path=0, symbol_len=int(symbol_len), n_branches=B,
                         n_frames=F, n_samples=n, x_frame_stride=B * n, x_branch_stride=n, out_stride=n_out, store_mode=0,
                         reserved=0)
        L.check(L.lib().ofs_park_metric(C.byref(d), _ptr(x), _ptr(M), _ptr(P), _ptr(E), _stream()), "ofs_park_metric")
    return M, P, E


# ------------------------------------------------------------------------------------------- sync_aa
def aa_metric_reference(rx, half_len: int):
    """Reference-order [A][A] running sums (sync_aa.py:321-386, 458-493) -> P c128, R, M f64, valid u8, each [F, n]."""
    x, code, _ = to_device(rx)
    F, A, n = x.shape[0], x.shape[1], x.shape[2]
    dev = x.device
    P = torch.empty((F, n), dtype=torch.complex128, device=dev)
    R = torch.empty((F, n), dtype=torch.float64, device=dev)
    M = torch.empty((F, n), dtype=torch.float64, device=dev)
    valid = torch.empty((F, n), dtype=torch.uint8, device=dev)
    L.check(L.lib().ofs_aa_metric_reference(_ptr(x), code, C.c_int64(F), int(A), C.c_int64(n), int(half_len), _ptr(P), _ptr(R),
                                            _ptr(M), _ptr(valid), _stream()), "ofs_aa_metric_reference")
    return P, R, M, valid


class AADetectPlan:
    """Fused antenna-array detector (ofs_aa_detect): captures [F, A, n] complex64 or int16-IQ [F, A, n, 2] on the device ->
    M float32 / P complex64 [F, n] + gate events (sync_aa.py:458-568), one pass over the samples."""

    def __init__(self, n_frames: int, n_antennas: int, n: int, half_len: int = 512, threshold: float = 0.15,
                 hysteresis: int = 128, sample_rate: float = 15.36e6, in_dtype: str = "c64", want_r: bool = False,
                 max_events: int = L.OFS_MAX_EVENTS):
        dev = _device()
        self.F, self.A, self.n, self.Lh = n_frames, n_antennas, n, half_len
        self.thr, self.hyst, self.fs = threshold, hysteresis, sample_rate
        self.code = {"c64": L.OFS_C64, "iq16": L.OFS_IQ16}[in_dtype]
        self.pitch = (n + 1) // 2 * 2
        self.Mb = torch.empty((n_frames, self.pitch), dtype=torch.float32, device=dev)
        self.Pb = torch.empty((n_frames, self.pitch), dtype=torch.complex64, device=dev)
        self.Rb = torch.empty((n_frames, self.pitch), dtype=torch.float32, device=dev) if want_r else None
        self.mstride = (n + 31) // 32
        self.mask = torch.zeros((n_frames, self.mstride), dtype=torch.int32, device=dev)
        self.cap = int(max_events)
        self.ev, self.cnt, self.records = _event_buffers(n_frames, dev, want_flat=True, cap=self.cap)   # records: what a rank gathers
        self.M, self.P = self.Mb[:, :n], self.Pb[:, :n]
        self.R = None if self.Rb is None else self.Rb[:, :n]

    def run(self, x: torch.Tensor) -> None:
        assert x.is_cuda and x.is_contiguous() and x.shape[0] == self.F and x.shape[1] == self.A and x.shape[2] == self.n
        L.check(L.lib().ofs_aa_detect(_ptr(x), self.code, C.c_int64(self.F), int(self.A), C.c_int64(self.n), C.c_int64(self.A * self.n),
                                      C.c_int64(self.n), int(self.Lh), C.c_double(self.thr), int(self.hyst), C.c_double(self.fs),
                                      _ptr(self.Mb), _ptr(self.Pb), _ptr(self.Rb), C.c_int64(self.pitch), _ptr(self.mask),
                                      C.c_int64(self.mstride), _ptr(self.ev), _ptr(self.cnt), int(self.cap), _stream()), "ofs_aa_detect")

    def events(self, overflow: str = "raise") -> list[np.ndarray]:
        """Events of the last run.  A capture with more gates than the plan's max_events slots raises EventOverflow
        (overflow="truncate": the first max_events, the counts in self.cnt stay true)."""
        return _events_to_numpy(self.ev, self.cnt, self.cap, overflow)


# ------------------------------------------------------------------------------------------- minn_rtl
def minn_rtl_metric(rx, quarter_len: int, smooth_shift: int, threshold_value: int, frac_bits: int) -> dict:
    """minn_rtl.py:583-733 (float mirror) -> dict of 8 tensors [F, n]."""
    x, code, _ = to_device(rx)
    if code == L.OFS_IQ16:
        raise TypeError("minn_rtl_metric takes complex input; use minn_rtl_int for int16 IQ")
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    dev = x.device
    f = lambda: torch.empty((F, n), dtype=torch.float64, device=dev)
    u = lambda: torch.empty((F, n), dtype=torch.uint8, device=dev)
    d = dict(corr_total=f(), corr_positive=f(), smooth_metric=f(), energy_total=f(), corr_scaled=f(), energy_scaled=f(),
             metric_valid=u(), above_threshold=u())
    rc = L.lib().ofs_minn_rtl_metric(_ptr(x), code, C.c_int64(F), int(B), C.c_int64(n), int(quarter_len), int(smooth_shift),
                                     int(threshold_value), int(frac_bits), *[_ptr(d[k]) for k in d], _stream())
    L.check(rc, "ofs_minn_rtl_metric")
    return d


def minn_rtl_int(iq, quarter_len: int, smooth_shift: int, threshold_value: int, frac_bits: int, lag_extra: int = 0) -> dict:
    """Integer RTL datapath (ref/minn_antenna_path.sv, ref/minn_preamble_detector.sv:247-325) -> int64 tensors [F, n]."""
    x, code, _ = to_device(iq)
    if code != L.OFS_IQ16:
        raise TypeError("minn_rtl_int takes int16 IQ input with a trailing axis of 2")
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    dev = x.device
    f = lambda: torch.empty((F, n), dtype=torch.int64, device=dev)
    u = lambda: torch.empty((F, n), dtype=torch.uint8, device=dev)
    d = dict(corr_total=f(), corr_positive=f(), smooth_metric=f(), energy_total=f(), metric_valid=u(), above_threshold=u())
    rc = L.lib().ofs_minn_rtl_int(_ptr(x), C.c_int64(F), int(B), C.c_int64(n), int(quarter_len), int(smooth_shift),
                                  int(threshold_value), int(frac_bits), int(lag_extra), *[_ptr(d[k]) for k in d], _stream())
    L.check(rc, "ofs_minn_rtl_int")
    return d


class RtlIntPlan:
    """Integer RTL datapath + gate FSM (ofs_minn_rtl_int + ofs_minn_rtl_events) for repeated batches of int16-IQ streams
    [F, B, n, 2] with every buffer allocated once: the six state arrays (int64 / uint8 [F, n]), event slots and counts.
    run() only launches (window sums + antenna combining, parallel floor-shift smoother, threshold, gate FSM); events() reads back."""

    def __init__(self, n_frames: int, n_branches: int, n: int, quarter_len: int = 512, smooth_shift: int = 3,
                 threshold_value: int = 3276, frac_bits: int = 15, hysteresis: int = 2, timing_offset: int = 0,
                 lag_extra: int = 0, max_events: int = L.OFS_MAX_EVENTS):
        dev = _device()
        self.F, self.B, self.n = n_frames, n_branches, n
        self.args = (int(quarter_len), int(smooth_shift), int(threshold_value), int(frac_bits), int(lag_extra))
        self.hyst, self.toff = int(hysteresis), int(timing_offset)
        f = lambda: torch.empty((n_frames, n), dtype=torch.int64, device=dev)
        u = lambda: torch.empty((n_frames, n), dtype=torch.uint8, device=dev)
        self.state = dict(corr_total=f(), corr_positive=f(), smooth_metric=f(), energy_total=f(), metric_valid=u(), above_threshold=u())
        self.cap = int(max_events)
        self.ev, self.cnt, self.records = _event_buffers(n_frames, dev, want_flat=True, cap=self.cap)

    def run(self, iq: torch.Tensor) -> None:
        assert iq.is_cuda and iq.is_contiguous() and iq.dtype == torch.int16 and tuple(iq.shape) == (self.F, self.B, self.n, 2)
        d = self.state
        L.check(L.lib().ofs_minn_rtl_int(_ptr(iq), C.c_int64(self.F), int(self.B), C.c_int64(self.n), *self.args,
                                         *[_ptr(d[k]) for k in d], _stream()), "ofs_minn_rtl_int")
        L.check(L.lib().ofs_minn_rtl_events(_ptr(d["corr_positive"]), 1, _ptr(d["metric_valid"]), _ptr(d["above_threshold"]),
                                            C.c_int64(self.F), C.c_int64(self.n), C.c_int64(self.n), self.hyst, self.toff,
                                            _ptr(self.ev), _ptr(self.cnt), int(self.cap), _stream()), "ofs_minn_rtl_events")

    def events(self, overflow: str = "raise") -> list[np.ndarray]:
        return _events_to_numpy(self.ev, self.cnt, self.cap, overflow)


# ------------------------------------------------------------------------------------------- Zadoff-Chu
def zc_matched_filter(rx, ref, mode: int = 0, out_f64: bool | None = None, want_corr: bool = True):
    """FFT overlap-save matched filter (zc.py:115-126 mode 0, zc_v2.py:486-495 mode 1, raw sum mode 2)
    -> (corr complex [F, n+nr-1] or None with want_corr=False, |corr|)."""
    x, code, _ = to_device(rx)
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    if out_f64 is None:
        out_f64 = code == L.OFS_C128
    r = torch.as_tensor(np.ascontiguousarray(np.asarray(ref, dtype=np.complex128))).to(x.device)
    nr = r.numel()
    n_out = n + nr - 1
    cdt = torch.complex128 if out_f64 else torch.complex64
    rdt = torch.float64 if out_f64 else torch.float32
    corr = torch.empty((F, n_out), dtype=cdt, device=x.device) if want_corr else None
    mag = torch.empty((F, n_out), dtype=rdt, device=x.device)
    L.check(L.lib().ofs_zc_matched_filter(_ptr(x), code, C.c_int64(F), int(B), C.c_int64(n), _ptr(r), int(nr), int(mode),
                                          int(out_f64), _ptr(corr), _ptr(mag), C.c_int64(n_out), _stream()), "ofs_zc_matched_filter")
    return corr, mag


def zc_v2_detect(rx, ref, *, normalize: bool = True, window: int = 2048, thresh_value: int = 64, frac_bits: int = 15,
                 min_corr_mag: float = 0.3, hysteresis: int = 256, max_events: int | None = None):
    """zc_v2.detect_zc_preamble (zc_v2.py:456-516) for a batch of complex64 / int16-IQ captures on the float32 path
    (ofs_zc_v2_detect: matched filter writing |corr| only -> running-sum threshold bitmask -> gate FSM).
    -> (events per capture, corr_mag float32 [F, n + nr - 1])."""
    x, code, _ = to_device(rx)
    if code == L.OFS_C128:
        raise TypeError("zc_v2_detect is the float32 pipeline: pass complex64 or int16 IQ (the drop-in zc_v2 module serves complex128)")
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    r = torch.as_tensor(np.ascontiguousarray(np.asarray(ref, dtype=np.complex128))).to(x.device)
    nr = r.numel()
    n_out = n + nr - 1
    mag = torch.empty((F, n_out), dtype=torch.float32, device=x.device)
    mstride = (n_out + 31) // 32
    mask = torch.empty((F, mstride), dtype=torch.int32, device=x.device)

    def launch(ev, cnt, cap):
        L.check(L.lib().ofs_zc_v2_detect(_ptr(x), code, C.c_int64(F), int(B), C.c_int64(n), _ptr(r), int(nr), int(bool(normalize)),
                                         int(window), int(thresh_value), int(frac_bits), C.c_double(min_corr_mag), int(hysteresis),
                                         _ptr(mag), C.c_int64(n_out), _ptr(mask), C.c_int64(mstride), _ptr(ev), _ptr(cnt), int(cap),
                                         _stream()), "ofs_zc_v2_detect")
    return _with_event_buffers(F, x.device, launch, max_events), mag


class ZCDetectPlan:
    """zc_v2.detect_zc_preamble for repeated batches of captures [F, B, n] complex64 / int16 IQ (ofs_zc_v2_detect) with all
    buffers allocated once: |corr| rows, threshold bitmask, event slots + counts.  run() only launches; events() reads back."""

    def __init__(self, n_frames: int, n_branches: int, n: int, ref, *, normalize: bool = True, window: int = 2048,
                 thresh_value: int = 64, frac_bits: int = 15, min_corr_mag: float = 0.3, hysteresis: int = 256,
                 in_dtype: str = "c64", max_events: int = L.OFS_MAX_EVENTS):
        dev = _device()
        self.F, self.B, self.n = n_frames, n_branches, n
        self.code = {"c64": L.OFS_C64, "iq16": L.OFS_IQ16}[in_dtype]
        self.ref = torch.as_tensor(np.ascontiguousarray(np.asarray(ref, dtype=np.complex128))).to(dev)
        self.nr = self.ref.numel()
        self.n_out = n + self.nr - 1
        self.args = (int(bool(normalize)), int(window), int(thresh_value), int(frac_bits), C.c_double(min_corr_mag), int(hysteresis))
        self.mag = torch.empty((n_frames, self.n_out), dtype=torch.float32, device=dev)
        self.mstride = (self.n_out + 31) // 32
        self.mask = torch.empty((n_frames, self.mstride), dtype=torch.int32, device=dev)
        self.cap = int(max_events)
        self.ev, self.cnt, self.records = _event_buffers(n_frames, dev, want_flat=True, cap=self.cap)

    def run(self, x: torch.Tensor) -> None:
        assert x.is_cuda and x.is_contiguous() and x.shape[0] == self.F and x.shape[1] == self.B and x.shape[2] == self.n
        L.check(L.lib().ofs_zc_v2_detect(_ptr(x), self.code, C.c_int64(self.F), int(self.B), C.c_int64(self.n), _ptr(self.ref),
                                         int(self.nr), *self.args, _ptr(self.mag), C.c_int64(self.n_out), _ptr(self.mask),
                                         C.c_int64(self.mstride), _ptr(self.ev), _ptr(self.cnt), int(self.cap), _stream()),
                "ofs_zc_v2_detect")

    def events(self, overflow: str = "raise") -> list[np.ndarray]:
        return _events_to_numpy(self.ev, self.cnt, self.cap, overflow)


def zc_normalize(corr, rx, reference):
    """zc_v2.normalize_correlation (zc_v2.py:257-271) of a caller-supplied correlation: corr (..., n + nr - 1) complex,
    rx (frames, n) or (n,) one branch -> corr / (||ref|| * sqrt(max(sliding energy, 1e-12))), same dtype as corr."""
    x, code, _ = to_device(rx)
    F, B, n = x.shape
    if B != 1:
        raise ValueError("normalize_correlation takes one branch")
    c = corr if isinstance(corr, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(corr))
    if c.dtype not in (torch.complex64, torch.complex128):
        c = c.to(torch.complex128)
    ref = np.asarray(reference)
    nr = ref.size
    c = c.to(x.device).reshape(F, -1).contiguous()
    if c.shape[1] != n + nr - 1:
        raise ValueError(f"corr has {c.shape[1]} samples per row, expected n + len(reference) - 1 = {n + nr - 1}")
    o = torch.empty_like(c)
    ref_norm = float(np.sqrt(np.sum(np.abs(ref) ** 2)))
    L.check(L.lib().ofs_zc_normalize(_ptr(c), _ptr(x), code, C.c_int64(F), C.c_int64(n), int(nr), C.c_double(ref_norm),
                                     int(c.dtype == torch.complex128), _ptr(o), C.c_int64(c.shape[1]), _stream()), "ofs_zc_normalize")
    return o


def zc_freq_metric(rx, bin_indices, template_bins, template_energy: float, n_fft: int = 2048, cp: int = 512,
                   out_f64: bool | None = None, fast: bool | str = False) -> torch.Tensor:
    """zc_freq.py:62-99 -> metric [F, n - (n_fft+cp) + 1].  Default: the float64-prefix kernel (1e-11).  For complex64
    captures: fast="fft" the FFT form (ofs_zc_freq_metric_fft: two overlap-save filters + the energy recurrence,
    float32 metric within 1e-4 of its maximum, n_fft <= 2048, any number of branches); single-branch only: fast="f32" the packed-fp32 sliding-DFT kernel
    (ofs_zc_freq_metric_f32, same tolerance, any n_fft multiple of 32); fast=True / "tc" the tensor-core kernel of the
    correlator bank (fp16 operands, 5e-3)."""
    x, code, _ = to_device(rx)
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    if out_f64 is None:
        out_f64 = code == L.OFS_C128
    n_off = n - (n_fft + cp) + 1
    if n_off <= 0:
        raise ValueError("Received stream is shorter than a single OFDM symbol.")     # zc_freq.py:76-78
    if fast:
        if fast == "fft":
            if code == L.OFS_C128 or out_f64:
                raise L.OfsError("zc_freq_metric(fast='fft') takes complex64 or int16-IQ captures and returns float32")
        elif code != L.OFS_C64 or out_f64 or B != 1:
            raise L.OfsError("zc_freq_metric(fast=...) takes complex64 single-branch captures (fast='fft': any branch count, int16 IQ too) and returns float32")
        k = np.mod(np.asarray(bin_indices, dtype=np.int64), n_fft).astype(np.int32)
        bins = torch.as_tensor(k).to(x.device)
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(template_bins, dtype=np.complex64))).to(x.device)
        out = torch.empty((F, n_off), dtype=torch.float32, device=x.device)
        if fast == "fft":
            L.check(L.lib().ofs_zc_freq_metric_fft(_ptr(x), code, C.c_int64(F), int(B), C.c_int64(n), int(n_fft), int(cp), _ptr(bins), _ptr(t),
                                                   int(k.size), C.c_double(float(template_energy)), _ptr(out), C.c_int64(n_off),
                                                   _stream()), "ofs_zc_freq_metric_fft")
            return out
        fn, name = (L.lib().ofs_zc_freq_metric_f32, "ofs_zc_freq_metric_f32") if fast == "f32" else \
                   (L.lib().ofs_zc_freq_metric_fast, "ofs_zc_freq_metric_fast")
        L.check(fn(_ptr(x), C.c_int64(F), C.c_int64(n), int(n_fft), int(cp), _ptr(bins), _ptr(t), int(k.size),
                   C.c_double(float(template_energy)), _ptr(out), C.c_int64(n_off), _stream()), name)
        return out
    # positions = (N/2 + idx) % N index the fftshifted spectrum == plain DFT bin idx mod N
    k = np.mod(np.asarray(bin_indices, dtype=np.int64), n_fft).astype(np.int32)
    bins = torch.as_tensor(k).to(x.device)
    t = torch.as_tensor(np.ascontiguousarray(np.asarray(template_bins, dtype=np.complex128))).to(x.device)
    out = torch.empty((F, n_off), dtype=torch.float64 if out_f64 else torch.float32, device=x.device)
    L.check(L.lib().ofs_zc_freq_metric(_ptr(x), code, C.c_int64(F), int(B), C.c_int64(n), int(n_fft), int(cp), _ptr(bins), _ptr(t),
                                       int(k.size), C.c_double(float(template_energy)), int(out_f64), _ptr(out), C.c_int64(n_off),
                                       _stream()), "ofs_zc_freq_metric")
    return out


def zc_bank(rx, bin_indices, templates, n_fft: int = 2048, cp: int = 512):
    """Tensor-core correlator bank: max over offsets of the zc_freq metric for every template row.
    rx: (frames, n) or (n,) complex64; templates: (n_roots <= 128, nbins <= 64) complex.
    -> (best_metric float32 [F, R], best_offset int32 [F, R])."""
    t = rx if isinstance(rx, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(rx))
    if t.dim() == 1:
        t = t[None]
    x = t.to(torch.complex64).to(_device()).contiguous()
    F, n = x.shape
    if n - (n_fft + cp) + 1 <= 0:
        raise ValueError("Received stream is shorter than a single OFDM symbol.")
    k = np.mod(np.asarray(bin_indices, dtype=np.int64), n_fft).astype(np.int32)
    bins = torch.as_tensor(k).to(x.device)
    T = torch.as_tensor(np.ascontiguousarray(np.asarray(templates, dtype=np.complex64))).to(x.device)
    R, nb = T.shape
    bm = torch.zeros((F, R), dtype=torch.float32, device=x.device)
    bo = torch.zeros((F, R), dtype=torch.int32, device=x.device)
    L.check(L.lib().ofs_zc_bank(_ptr(x), C.c_int64(F), C.c_int64(n), int(n_fft), int(cp), _ptr(bins), _ptr(T), int(nb), int(R),
                                _ptr(bm), _ptr(bo), _stream()), "ofs_zc_bank")
    return bm, bo


# ------------------------------------------------------------------------------------------- channel / CFO
def channel_apply(tx, taps=None, *, row_of_stream=None, unit_noise=None, snr_db=None, cfo_hz=None, fs: float = 30.72e6,
                  full_scale=None, bits: int = 12, want_iq: bool = False):
    """channel.apply_channel -> core.apply_cfo -> sync_aa.quantize_adc on the device (ofs_channel_apply).
    tx: [rows, n_tx] complex64 / complex128; taps: 1-D complex or None; unit_noise: [streams, n_out] unit-variance complex
    normal samples per component (or None); snr_db / cfo_hz / full_scale: per-stream arrays (or None).
    -> (out [streams, n_out] complex, out_iq int16 [streams, n_out, 2] or None)."""
    t = tx if isinstance(tx, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(tx))
    if t.dim() == 1:
        t = t[None]
    if t.dtype not in (torch.complex64, torch.complex128):
        t = t.to(torch.complex128)
    dev = _device()
    t = t.to(dev).contiguous()
    code = L.OFS_C64 if t.dtype == torch.complex64 else L.OFS_C128
    rows, n_tx = t.shape
    n_taps = 0
    tp = None
    if taps is not None:
        tp = torch.as_tensor(np.ascontiguousarray(np.asarray(taps, dtype=np.complex128).reshape(-1))).to(dev)
        n_taps = tp.numel()
    n_out = n_tx + n_taps - 1 if tp is not None else n_tx
    ros = None
    S = rows
    if row_of_stream is not None:
        ros = torch.as_tensor(np.asarray(row_of_stream, dtype=np.int32)).to(dev)
        S = ros.numel()
    dv = lambda a: None if a is None else torch.as_tensor(np.broadcast_to(np.asarray(a, dtype=np.float64), (S,)).copy()).to(dev)
    snr, cfo, fsc = dv(snr_db), dv(cfo_hz), dv(full_scale)
    nz = None
    if unit_noise is not None:
        nz = unit_noise if isinstance(unit_noise, torch.Tensor) else torch.as_tensor(np.array(unit_noise, order="C"))
        nz = nz.to(device=dev, dtype=t.dtype).reshape(S, -1).contiguous()
    out = torch.empty((S, n_out), dtype=t.dtype, device=dev)
    iq = torch.empty((S, n_out, 2), dtype=torch.int16, device=dev) if want_iq else None
    ws = torch.empty((rows, n_out), dtype=t.dtype, device=dev)
    pw = torch.zeros(rows, dtype=torch.float64, device=dev)
    L.check(L.lib().ofs_channel_apply(_ptr(t), code, C.c_int64(rows), C.c_int64(n_tx), _ptr(tp), int(n_taps), C.c_int64(S), _ptr(ros),
                                      _ptr(nz), C.c_int64(0 if nz is None else nz.shape[1]), _ptr(snr), _ptr(cfo), C.c_double(fs),
                                      _ptr(fsc), int(bits), _ptr(out), _ptr(iq), C.c_int64(n_out), _ptr(ws), _ptr(pw), _stream()),
            "ofs_channel_apply")
    return out, iq


def wire_pack(iq, fmt: str) -> torch.Tensor:
    """int16 IQ -> the RTL side's 12-bit words (ofs_wire_pack).  fmt "hex24": iq [n, 2] -> int32[n] words {Re, Im}
    (docs/preamble_test_vector.hex); fmt "axis48": iq [2, n, 2] -> int64[n] words {ch1_q, ch1_i, ch0_q, ch0_i}
    (ref/test_minn_preamble_detector.py:41-47)."""
    t = iq if isinstance(iq, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(iq))
    if t.dtype != torch.int16 or t.shape[-1] != 2:
        raise TypeError("wire_pack takes int16 IQ with a trailing axis of 2")
    t = t.to(_device()).contiguous()
    if fmt == "hex24":
        if t.dim() != 2:
            raise ValueError("hex24 packs one channel: iq [n, 2]")
        n, code, w = t.shape[0], L.OFS_WIRE_HEX24, torch.empty(t.shape[0], dtype=torch.int32, device=t.device)
    elif fmt == "axis48":
        if t.dim() != 3 or t.shape[0] != 2:
            raise ValueError("axis48 packs two channels: iq [2, n, 2]")
        n, code, w = t.shape[1], L.OFS_WIRE_AXIS48, torch.empty(t.shape[1], dtype=torch.int64, device=t.device)
    else:
        raise ValueError(f"unknown wire format {fmt!r}")
    if n:
        L.check(L.lib().ofs_wire_pack(_ptr(t), C.c_int64(n), int(code), _ptr(w), _stream()), "ofs_wire_pack")
    return w


def wire_unpack(words, fmt: str) -> torch.Tensor:
    """The inverse of wire_pack (ofs_wire_unpack): words -> int16 IQ [n, 2] (hex24) or [2, n, 2] (axis48), sign-extended."""
    dt = {"hex24": torch.int32, "axis48": torch.int64}.get(fmt)
    if dt is None:
        raise ValueError(f"unknown wire format {fmt!r}")
    w = words if isinstance(words, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(np.asarray(words).astype(np.int64)))
    w = w.to(device=_device(), dtype=dt).reshape(-1).contiguous()
    n = w.numel()
    iq = torch.empty((n, 2) if fmt == "hex24" else (2, n, 2), dtype=torch.int16, device=w.device)
    code = L.OFS_WIRE_HEX24 if fmt == "hex24" else L.OFS_WIRE_AXIS48
    if n:
        L.check(L.lib().ofs_wire_unpack(_ptr(w), C.c_int64(n), int(code), _ptr(iq), _stream()), "ofs_wire_unpack")
    return iq


def cp_cfo(rx, starts, n_fft: int = 2048, cp_len: int = 512, fs: float = 30.72e6, mode: str = "plain", span: int | None = None,
           win_len: int | None = None):
    """core.estimate_cfo_from_cp / _robust / _peak_with_index (core.py:179-308) for a batch of frames.
    -> (cfo_hz float64[F], best_d int64[F], P complex128[F]) tensors."""
    x, code, _ = to_device(rx)
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    st = torch.as_tensor(np.broadcast_to(np.asarray(starts, dtype=np.int64), (F,)).copy()).to(x.device)
    m = {"plain": 0, "robust": 1, "peak": 2}[mode]
    sp = cp_len // 2 if span is None else int(max(0, span))                     # core.py:217,246
    wl = cp_len // 2 if win_len is None else int(max(1, win_len))               # core.py:218
    cfo = torch.empty(F, dtype=torch.float64, device=x.device)
    bd = torch.empty(F, dtype=torch.int64, device=x.device)
    P = torch.empty(F, dtype=torch.complex128, device=x.device)
    L.check(L.lib().ofs_cp_cfo(_ptr(x), code, C.c_int64(F), int(B), C.c_int64(n), C.c_int64(B * n), C.c_int64(n), _ptr(st), int(n_fft),
                               int(cp_len), int(sp), int(wl), int(m), C.c_double(fs), _ptr(cfo), _ptr(bd), _ptr(P), _stream()), "ofs_cp_cfo")
    return cfo, bd, P


def rx_chain(rx, pilot_cp_start, cfo_hz, pilot_used, data_used, *, fs: float = 30.72e6, n_fft: int = 2048, cp_len: int = 512,
             used_indices=None):
    """The receive chain after the detector, batched (ofs_rx_chain): CFO correction, pilot FFT -> LS estimate -> phase-slope
    timing, data FFT -> equalise -> gain alignment -> EVM (sc.py:286-309 / core.py:171-176, 339-370, 443-469).
    rx: (frames, branches, n); pilot_cp_start, cfo_hz: per frame; pilot_used: [n_used]; data_used: [n_used] or [frames, n_used].
    -> dict(h_est [F, n_used] c128, xhat [F, n_used] c128, evm_rms, evm_db, slope, sto, gain)."""
    x, code, _ = to_device(rx)
    F, B, n = x.shape[0], x.shape[1], x.shape[2]
    dev = x.device
    if used_indices is None:
        half = len(np.asarray(pilot_used)) // 2
        used_indices = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))      # core.centered_subcarrier_indices
    k = np.asarray(used_indices, dtype=np.int64)
    nu = k.size
    bins = torch.as_tensor(np.mod(k, n_fft).astype(np.int32)).to(dev)
    kf = torch.as_tensor(k.astype(np.float64)).to(dev)
    st = torch.as_tensor(np.broadcast_to(np.asarray(pilot_cp_start, dtype=np.int64), (F,)).copy()).to(dev)
    cf = torch.as_tensor(np.broadcast_to(np.asarray(cfo_hz, dtype=np.float64), (F,)).copy()).to(dev)
    pu = torch.as_tensor(np.ascontiguousarray(np.asarray(pilot_used, dtype=np.complex128))).to(dev)
    du = np.asarray(data_used, dtype=np.complex128)
    dstride = 0 if du.ndim == 1 else nu
    dd = torch.as_tensor(np.ascontiguousarray(du)).to(dev)
    h = torch.empty((F, nu), dtype=torch.complex128, device=dev)
    xh = torch.empty((F, nu), dtype=torch.complex128, device=dev)
    sc = torch.zeros((F, 8), dtype=torch.float64, device=dev)
    L.check(L.lib().ofs_rx_chain(_ptr(x), code, C.c_int64(F), int(B), C.c_int64(n), C.c_int64(B * n), C.c_int64(n), _ptr(st), _ptr(cf),
                                 C.c_double(fs), int(n_fft), int(cp_len), _ptr(bins), _ptr(kf), int(nu), _ptr(pu), _ptr(dd),
                                 C.c_int64(dstride), _ptr(h), _ptr(xh), _ptr(sc), _stream()), "ofs_rx_chain")
    return dict(h_est=h, xhat=xh, evm_rms=sc[:, 0], evm_db=sc[:, 1], slope=sc[:, 2], sto=sc[:, 3],
                gain=torch.complex(sc[:, 4], sc[:, 5]), valid=sc[:, 7])
