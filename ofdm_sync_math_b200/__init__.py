"""ofdm_sync_math_b200 -- B200-native OFDM synchronisation engine (hot path of amcolex/ofdm-sync-math).

Drop-in modules (same function names and return shapes as the reference's scripts):
    sc, minn, park, combined_sc_min, zc, zc_v2, zc_freq, sync_aa, minn_rtl
Batched device API: engine (torch tensors), dist (sharding over GPUs + NCCL gather of records).
All array work runs in hand-written sm_100a kernels in libofdmsync.so (C ABI: include/ofdmsync.h);
there is no CPU fallback -- importing the compute entry points without the library raises.
"""
from . import _lib  # noqa: F401
from ._lib import OfsError, build  # noqa: F401

__all__ = ["sc", "minn", "park", "combined_sc_min", "zc", "zc_v2", "zc_freq", "sync_aa", "minn_rtl", "engine", "dist",
           "patch_reference", "OfsError", "build"]

HOT_FUNCTIONS = {
    "sc": ["sc_streaming_metric", "find_plateau_end_from_metric"],
    "minn": ["minn_streaming_metric", "minn_streaming_metric_parameterized", "_trailing_average", "find_minn_peak"],
    "park": ["park_streaming_metric"],
    "combined_sc_min": ["minn_streaming_metric", "schmidl_cox_streaming_metric", "_trailing_average",
                        "_streaming_peak_detector", "find_minn_peak"],
    "zc_v2": ["matched_filter_correlation", "normalize_correlation", "zc_streaming_detection", "detect_zc_peaks",
              "detect_zc_preamble"],
    "zc_freq": ["compute_frequency_metric"],
    "sync_aa": ["aa_detect_streaming"],
    "minn_rtl": ["minn_rtl_streaming_metric", "detect_minn_rtl"],
}


# the stages either side of the hot path and the sweep drivers, batched on the device.  Opt-in
# (patch_reference(module, sweeps=True)): these replace whole experiment loops, not just kernels.
SWEEP_FUNCTIONS = {
    "channel": ["apply_channel"],
    "core": ["apply_cfo", "estimate_cfo_from_cp", "estimate_cfo_from_cp_robust", "estimate_cfo_from_cp_peak",
             "estimate_cfo_from_cp_peak_with_index", "find_cp_start_via_corr"],
    "sync_aa": ["quantize_adc", "apply_cfo", "run_single_test", "run_grid_test"],
    "minn": ["compare_block_lengths"],
    "minn_rtl": ["compare_q_values"],
}


def patch_reference(module, name: str | None = None, sweeps: bool = False) -> list[str]:
    """Replace the hot functions of an imported reference module (e.g. `import sc`) with this engine's,
    so the reference's own run_simulation()/plots run unchanged on top of the CUDA kernels.  sweeps=True also swaps the
    impairment chain / CFO estimators / sweep drivers of SWEEP_FUNCTIONS (sync_aa.main, plot_snr_sweep, plot_q_comparison then
    run their cases as device batches)."""
    import importlib
    name = name or module.__name__.split(".")[-1]
    ours = importlib.import_module(f"{__name__}.{name}")
    done = []
    for fn in HOT_FUNCTIONS.get(name, []) + (SWEEP_FUNCTIONS.get(name, []) if sweeps else []):
        if hasattr(module, fn):
            setattr(module, fn, getattr(ours, fn))
            done.append(fn)
    if sweeps:
        # the scripts pull the channel / CFO helpers into their own namespace (`from channel import apply_channel`): swap those
        # bindings too, but only names that really came from channel.py / core.py
        for src in ("channel", "core"):
            if src == name:
                continue
            helper = importlib.import_module(f"{__name__}.{src}")
            for fn in SWEEP_FUNCTIONS[src]:
                cur = getattr(module, fn, None)
                if cur is not None and fn not in done and getattr(cur, "__module__", None) == src:
                    setattr(module, fn, getattr(helper, fn))
                    done.append(fn)
    return done
