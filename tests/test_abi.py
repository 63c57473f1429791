"""CPU checks of the C-ABI boundary: libofdmsync.so builds/loads without a GPU, exports every symbol that
include/ofdmsync.h declares, struct layouts match the ctypes mirror, and the package refuses to compute
without a CUDA device (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    txt = (ROOT / "include" / "ofdmsync.h").read_text()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ofs_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from ofdm_sync_math_b200 import _lib
    if not _lib.LIB_PATH.exists():
        _lib.build()
    lib = C.CDLL(str(_lib.LIB_PATH))
    syms = _declared_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, f"declared in ofdmsync.h but not exported: {missing}"
    assert lib.ofs_version() == 2
    assert lib.ofs_chunk_len() == 256


def test_struct_layouts():
    from ofdm_sync_math_b200 import _lib, engine
    assert C.sizeof(_lib.MetricDesc) == 72
    assert C.sizeof(_lib.Rows) == 40
    assert C.sizeof(_lib.Event) == 72
    assert C.sizeof(_lib.SyncRecord) == 40 and C.sizeof(_lib.SyncParams) == 32
    assert engine._EVENT_NP.itemsize == 72 and engine._REC_NP.itemsize == 40


def test_out_len_and_stripe_predicate_without_gpu():
    from ofdm_sync_math_b200 import _lib
    lib = _lib.lib()
    d = _lib.MetricDesc(kind=_lib.OFS_SC, in_dtype=_lib.OFS_C64, out_f64=0, path=0, symbol_len=2048, n_branches=1,
                        n_frames=4, n_samples=10000, x_frame_stride=10000, x_branch_stride=10000, out_stride=10000)
    assert lib.ofs_metric_out_len(C.byref(d)) == 10000 - 2048 + 1
    assert lib.ofs_metric_stripe_ok(C.byref(d), None, None) == 1
    d.symbol_len = 1000
    assert lib.ofs_metric_stripe_ok(C.byref(d), None, None) == 0
    d.kind = _lib.OFS_AA; d.symbol_len = 512
    assert lib.ofs_metric_out_len(C.byref(d)) == 10000
    d.kind = _lib.OFS_SC; d.symbol_len = 2048; d.n_samples = 100
    assert lib.ofs_metric_out_len(C.byref(d)) == 0
    # invalid descriptor -> negative code + message, nothing thrown
    d.kind = 99
    rc = lib.ofs_metric(C.byref(d), None, None, None, None, None, C.c_int64(0), None)
    assert rc == -1 and b"kind" in lib.ofs_last_error_string()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ofdm_sync_math_b200 import OfsError, sc
    with pytest.raises(OfsError):
        sc.sc_streaming_metric(np.zeros(5000, complex))
    # short inputs need no compute and follow the reference's empty-array rule (sc.py:50-52)
    M, P, R = sc.sc_streaming_metric(np.zeros(100, complex))
    assert M.size == 0 and P.size == 0 and P.dtype == complex and R.size == 0


def test_product_never_imports_oracle():
    pkg = ROOT / "ofdm_sync_math_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")):
        txt = f.read_text()
        assert "import oracle" not in txt and "from oracle" not in txt and "libofs_oracle" not in txt, f


def test_patch_reference_tables_and_imported_helpers():
    """patch_reference swaps the hot functions by name; with sweeps=True also the sweep drivers and the channel / core helpers a
    script imported into its own namespace (only bindings that really came from channel.py / core.py)."""
    import importlib
    import types
    import ofdm_sync_math_b200 as b200
    for table in (b200.HOT_FUNCTIONS, b200.SWEEP_FUNCTIONS):
        for name, fns in table.items():
            ours = importlib.import_module(f"ofdm_sync_math_b200.{name}")
            assert [f for f in fns if not hasattr(ours, f)] == [], name

    def stub(mod):
        f = lambda *a, **k: None
        f.__module__ = mod
        return f
    fake = types.ModuleType("minn")
    for fn, mod in (("minn_streaming_metric", "minn"), ("find_minn_peak", "minn"), ("compare_block_lengths", "minn"),
                    ("apply_channel", "channel"), ("apply_cfo", "core"), ("estimate_cfo_from_cp", "core"), ("load_measured_cir", "channel")):
        setattr(fake, fn, stub(mod))
    own_cfo = stub("minn")                                   # a script's own function that merely shares a helper's name stays
    fake.estimate_cfo_from_cp_robust = own_cfo
    keep = fake.compare_block_lengths
    assert sorted(b200.patch_reference(fake)) == ["find_minn_peak", "minn_streaming_metric"]
    assert fake.compare_block_lengths is keep
    done = b200.patch_reference(fake, sweeps=True)
    assert {"compare_block_lengths", "apply_channel", "apply_cfo", "estimate_cfo_from_cp"} <= set(done)
    from ofdm_sync_math_b200 import channel, core, minn
    assert fake.apply_channel is channel.apply_channel and fake.apply_cfo is core.apply_cfo
    assert fake.compare_block_lengths is minn.compare_block_lengths
    assert fake.estimate_cfo_from_cp_robust is own_cfo and "load_measured_cir" not in done


def test_bench_stdout_guard_keeps_one_json_line():
    """bench.StdoutToStderr: what a library writes to fd 1 during a multi-GPU run (NCCL's version banner) must not reach stdout,
    which carries exactly the JSON line."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = ("import os, bench\n"
            "g = bench.StdoutToStderr().start()\n"
            "os.write(1, b'NCCL version 2.28.9+cuda12.9\\n')\n"
            "print('python-level noise')\n"
            "g.stop()\n"
            "print('{\"ok\": 1}')\n")
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\n'
    assert "NCCL version" in r.stderr and "python-level noise" in r.stderr
