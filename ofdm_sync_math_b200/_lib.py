"""ctypes binding of libofdmsync.so (C ABI in include/ofdmsync.h).

The CUDA library is the product: if it is missing or does not load this module raises -- there is
no CPU fallback anywhere in the package (and nothing here imports oracle/).
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("OFS_LIB") or (_PKG / "libofdmsync.so"))   # OFS_LIB: A/B builds during kernel work

OFS_C64, OFS_C128, OFS_IQ16 = 0, 1, 2
OFS_SC, OFS_SC_BOTH, OFS_MINN, OFS_AA = 0, 1, 2, 3
OFS_PATH_AUTO, OFS_PATH_STRIPE, OFS_PATH_TILE, OFS_PATH_ARRAY = 0, 1, 2, 3
OFS_WIRE_HEX24, OFS_WIRE_AXIS48 = 0, 1
OFS_MAX_EVENTS = 64

i32, i64, f64, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p


class MetricDesc(C.Structure):
    _fields_ = [("kind", i32), ("in_dtype", i32), ("out_f64", i32), ("path", i32), ("symbol_len", i32),
                ("n_branches", i32), ("n_frames", i64), ("n_samples", i64), ("x_frame_stride", i64),
                ("x_branch_stride", i64), ("out_stride", i64), ("store_mode", i32), ("reserved", i32)]


class Rows(C.Structure):
    _fields_ = [("data", vp), ("f64", i32), ("reserved", i32), ("n_rows", i64), ("n", i64), ("stride", i64)]


class Event(C.Structure):
    _fields_ = [("peak_index", i64), ("gate_start", i64), ("gate_end", i64), ("aux", i64), ("value", f64),
                ("p_re", f64), ("p_im", f64), ("cfo", f64), ("closed", i32), ("reserved", i32)]


class SyncRecord(C.Structure):
    _fields_ = [("timing", i64), ("coarse", i64), ("metric", C.c_float), ("p_re", C.c_float),
                ("p_im", C.c_float), ("cfo", C.c_float), ("status", i32), ("reserved", i32)]


class SyncParams(C.Structure):
    _fields_ = [("cp_len", i32), ("smooth_win", i32), ("sc_delta", i32), ("exact", i32), ("gate_threshold", f64),
                ("exact_band", f64)]


OFS_ST_EXACT, OFS_ST_CHANGED, OFS_ST_UNRESOLVED = 1, 2, 4
OFS_EXACT_BAND = 1e-5


class OfsError(RuntimeError):
    pass


_lib = None


def build(verbose: bool = False) -> Path:
    """Compile csrc/*.cu for sm_100a into libofdmsync.so (nvcc cross-compiles without a GPU)."""
    import subprocess
    out = subprocess.run(["make", "-C", str(_PKG / "csrc"), "-j", str(os.cpu_count() or 4)],
                         capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout[-4000:], out.stderr[-4000:])
    if out.returncode != 0:
        raise OfsError("building libofdmsync.so failed")
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise OfsError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.ofs_last_error_string.restype = C.c_char_p
        _lib.ofs_metric_out_len.restype = i64
        _lib.ofs_launch_count.restype = i64
        _lib.ofs_host_alloc.restype = vp
        _lib.ofs_host_alloc.argtypes = [C.c_size_t]
        _lib.ofs_host_free.argtypes = [vp]
        _lib.ofs_ctx_destroy.argtypes = [vp]
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().ofs_last_error_string().decode(errors="replace")
        raise OfsError(f"{what or 'libofdmsync'} failed (rc={rc}): {msg}")


def launch_count() -> int:
    return int(lib().ofs_launch_count())
