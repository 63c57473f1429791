#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (amcolex/ofdm-sync-math).

TEST INFRASTRUCTURE ONLY.  Runs in the build container (where /root/reference is mounted);
the produced fixtures are committed so that the GPU box (which has no /root/reference) can
pin both the oracle (oracle/) and the CUDA path against real reference outputs.

How values are captured: the reference modules are imported unmodified behind a matplotlib
stub (oracle/refshim) and each script's own `run_simulation()` is executed; a `sys.setprofile`
hook grabs the local variables of `run_simulation` at return, so every array stored here was
computed by reference code (sc.py:159-368, minn.py:300-653, park.py:123-348,
combined_sc_min.py:272-580, zc.py:57-283, zc_v2.py:522-787, zc_freq.py:102-290,
minn_rtl.py:849-1100).  sync_aa / minn_rtl function-level cases call the reference functions
directly on inputs built with the reference's own builders.  SURVEY 8(f) fixtures: the impairment chain and the CP-CFO
estimators (channel_cfo), the receive chain (rxchain_*), the 135-case sync_aa grid (aa_grid), the block-length / Q
sweeps (sweeps) and the RTL testbench's AXIS word packer (wire_axis) -- all outputs of reference code, never of ours.

Usage:  python oracle/gen_golden.py [--ref /root/reference] [--out tests/golden] [--only sc,minn,...,aa_grid,sweeps,wire]
"""
from __future__ import annotations

import argparse
import os
import sys
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def _setup(ref: str) -> None:
    sys.path.insert(0, str(HERE / "refshim"))
    sys.path.insert(0, ref)
    sys.path.insert(0, str(Path(ref) / "ref"))
    os.chdir(tempfile.mkdtemp(prefix="ofs_golden_"))  # scripts mkdir plots/ under cwd


def _run_capture(module, fn_name: str, args: tuple) -> dict:
    """Run module.fn_name(*args) and return a copy of its locals at return."""
    captured: dict = {}
    target_code = getattr(module, fn_name).__code__

    def prof(frame, event, _arg):
        if event == "return" and frame.f_code is target_code:
            captured.update(frame.f_locals)

    sys.setprofile(prof)
    try:
        getattr(module, fn_name)(*args)
    finally:
        sys.setprofile(None)
    return captured


def _segments(mask: np.ndarray) -> np.ndarray:
    m = np.asarray(mask, dtype=bool)
    d = np.diff(np.concatenate(([0], m.astype(np.int8), [0])))
    return np.stack([np.flatnonzero(d == 1), np.flatnonzero(d == -1)], axis=1).astype(np.int64)


SCEN = [("cir1", "measured_channel"), (None, "flat_awgn")]


def gen_sc(out: Path) -> None:
    import sc
    for ch, sub in SCEN:
        loc = _run_capture(sc, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        np.savez_compressed(
            out / f"sc_{tag}.npz",
            rx=loc["rx_samples"], M=loc["M"], P=loc["P_sum"], R=loc["R_sum"],
            plateau_end=np.int64(loc["plateau_end"]), coarse_start=np.int64(loc["coarse_start"]),
            cp_len=np.int64(sc.CYCLIC_PREFIX), lookahead=np.int64(sc.CYCLIC_PREFIX // 4),
            smooth_win=np.int64(sc.SMOOTH_WIN), sc_delta=np.int64(sc.SC_DELTA),
            cfo_est_hz=np.float64(loc.get("cfo_est_hz", np.nan)),
        )
        print("sc", tag, int(loc["plateau_end"]), int(loc["coarse_start"]))


def gen_minn(out: Path) -> None:
    import minn
    for ch, sub in SCEN:
        loc = _run_capture(minn, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        np.savez_compressed(
            out / f"minn_{tag}.npz",
            rx=loc["rx_samples"], M=loc["M"], P=loc["P_sum"], R=loc["R_sum"],
            peak=np.int64(loc["peak_position"]), gate=_segments(loc["minn_gate_mask"]),
            Ms=loc["M_smooth"], smooth_win=np.int64(minn.SMOOTH_WIN),
            gate_threshold=np.float64(minn.MINN_GATE_THRESHOLD),
        )
        print("minn", tag, int(loc["peak_position"]), _segments(loc["minn_gate_mask"]).tolist())
    # parameterised symbol length on a short random input (minn.py:697-751)
    rng = np.random.default_rng(7)
    x = rng.standard_normal((2, 1500)) + 1j * rng.standard_normal((2, 1500))
    d = {"rx": x}
    for n in (64, 100, 256, 1024):
        M, P, R = minn.minn_streaming_metric_parameterized(x, n)
        d[f"M_{n}"], d[f"P_{n}"], d[f"R_{n}"] = M, P, R
    np.savez_compressed(out / "minn_param.npz", **d)


def gen_park(out: Path) -> None:
    import park
    for ch, sub in SCEN:
        loc = _run_capture(park, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        np.savez_compressed(
            out / f"park_{tag}.npz",
            rx=loc["rx_samples"], ds=loc["ds"], M=loc["M"], P=loc["P_sum"], E=loc["E_sum"],
            det_center=np.int64(loc["det_center"]), det_symbol_start=np.int64(loc["det_symbol_start"]),
        )
        print("park", tag, int(loc["det_center"]), int(loc["det_symbol_start"]))


def gen_combined(out: Path) -> None:
    import combined_sc_min as c
    for ch, sub in SCEN:
        loc = _run_capture(c, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        np.savez_compressed(
            out / f"combined_{tag}.npz",
            rx=loc["rx_samples"], M=loc["M"], P=loc["P_sum"], R=loc["R_sum"],
            M_sc=loc["M_sc"], P_sc=loc["P_sc"], R_sc=loc["R_sc"],
            sc_gate=_segments(loc["sc_gate_mask"]), sc_gate_span=np.asarray(loc["sc_gate_span"], dtype=np.int64),
            peak=np.int64(loc["peak_position"]), smooth_win=np.int64(c.SMOOTH_WIN),
            sc_gate_threshold=np.float64(c.SC_GATE_THRESHOLD),
        )
        print("combined", tag, int(loc["peak_position"]), loc["sc_gate_span"])


def gen_zc(out: Path) -> None:
    import zc
    for ch, sub in SCEN:
        loc = _run_capture(zc, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        np.savez_compressed(
            out / f"zc_{tag}.npz",
            rx=loc["rx_samples"], ref=loc["pss_reference"], corr=loc["combined_corr"],
            peak=np.int64(loc["peak_index"]), start=np.int64(loc["detected_start"]),
        )
        print("zc", tag, int(loc["peak_index"]), int(loc["detected_start"]))


def gen_zc_v2(out: Path) -> None:
    import zc_v2
    for ch, sub in SCEN:
        loc = _run_capture(zc_v2, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        res = loc["result"]
        st = res.state
        ev = np.array(
            [[e.peak_index, e.gate_start, e.gate_end, e.detected_start] for e in res.events], dtype=np.int64
        ).reshape(-1, 4)
        evv = np.array([e.peak_value for e in res.events], dtype=np.float64)
        np.savez_compressed(
            out / f"zc_v2_{tag}.npz",
            rx=loc["rx_samples"], ref=zc_v2.build_pss_symbol(include_cp=False),
            corr_mag=st.corr_mag, local_sum=st.local_sum, above=st.above_threshold,
            valid=st.metric_valid, gate=_segments(res.gate_mask), events=ev, event_values=evv,
            window_size=np.int64(zc_v2.CORR_WINDOW_SIZE), thresh_value=np.int64(zc_v2.THRESH_VALUE),
            thresh_frac_bits=np.int64(zc_v2.THRESH_FRAC_BITS), min_corr_mag=np.float64(zc_v2.MIN_CORR_MAG),
            hysteresis=np.int64(zc_v2.HYSTERESIS),
        )
        print("zc_v2", tag, ev.tolist(), evv.tolist())


def gen_zc_freq(out: Path) -> None:
    import zc_freq
    for ch, sub in SCEN:
        loc = _run_capture(zc_freq, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        np.savez_compressed(
            out / f"zc_freq_{tag}.npz",
            rx=loc["rx_samples"], metric=loc["metric"], peak=np.int64(loc["peak_index"]),
            bin_indices=loc["bin_positions"], template=loc["template_bins"],
            template_energy=np.float64(loc["template_energy"]),
        )
        print("zc_freq", tag, int(loc["peak_index"]))


def _rtl_state_dict(st, det) -> dict:
    ev = np.array(
        [[e.peak_index, e.detected_index, e.gate_segment[0], e.gate_segment[1]] for e in det.events],
        dtype=np.int64,
    ).reshape(-1, 4)
    return dict(
        corr_total=st.corr_total, corr_positive=st.corr_positive, smooth_metric=st.smooth_metric,
        energy_total=st.energy_total, corr_scaled=st.corr_scaled, energy_scaled=st.energy_scaled,
        metric_valid=st.metric_valid, above=st.above_threshold, events=ev,
        gate_segments=np.asarray(det.gate_segments, dtype=np.int64).reshape(-1, 2),
    )


def gen_minn_rtl(out: Path) -> None:
    import minn_rtl as mr
    for ch, sub in SCEN:
        loc = _run_capture(mr, "run_simulation", (ch, sub))
        tag = ch or "awgn"
        d = _rtl_state_dict(loc["metric_state"], loc["detection"])
        np.savez_compressed(
            out / f"minn_rtl_{tag}.npz", rx=loc["rx_samples"],
            smooth_shift=np.int64(mr.SMOOTH_SHIFT), threshold_value=np.int64(mr.THRESH_VALUE),
            threshold_frac_bits=np.int64(mr.THRESH_FRAC_BITS), quarter_len=np.int64(mr.PREAMBLE_Q),
            hysteresis=np.int64(mr.HYSTERESIS), timing_offset=np.int64(mr.TIMING_OFFSET), **d,
        )
        print("minn_rtl", tag, d["events"].tolist())

    # int12-valued, 2-antenna stimulus following the cocotb recipe
    # (ref/test_minn_preamble_detector.py:27-38,150-161,193-208) with a SEEDED data symbol
    # (the testbench leaves it unseeded, ref/ofdm.py:115-116).
    import ofdm
    params = ofdm.OFDMParameters(n_fft=2048, cp_len=512)
    preamble, _ = ofdm.generate_preamble(params=params)
    data_symbol, _ = ofdm.generate_qpsk_symbol(params=params, rng=np.random.default_rng(0))
    full = np.concatenate((np.zeros(256, complex), preamble, data_symbol, np.zeros(2048 + 512, complex)))
    rng = np.random.default_rng(0)

    def awgn(sig):
        p = np.mean(np.abs(sig) ** 2)
        npow = p / (10 ** (10.0 / 10))
        return sig + np.sqrt(npow / 2) * (rng.standard_normal(sig.shape) + 1j * rng.standard_normal(sig.shape))

    def quant(s, width=12):
        lo, hi = -(1 << (width - 1)), (1 << (width - 1)) - 1
        scale = (hi - 1) / np.max(np.abs(s))
        sc_ = s * scale
        return (np.clip(np.round(sc_.real), lo, hi).astype(np.int16),
                np.clip(np.round(sc_.imag), lo, hi).astype(np.int16))

    iq = np.zeros((2, full.size, 2), dtype=np.int16)
    for a in range(2):
        re, im = quant(awgn(full))
        iq[a, :, 0], iq[a, :, 1] = re, im
    rx = iq[..., 0].astype(np.float64) + 1j * iq[..., 1].astype(np.float64)
    kw = dict(smooth_shift=3, threshold_value=int(0.1 * (1 << 15)), threshold_frac_bits=15, quarter_len=512)
    st = mr.minn_rtl_streaming_metric(rx, **kw)
    det = mr.detect_minn_rtl(st, hysteresis=2, timing_offset=0)
    np.savez_compressed(
        out / "minn_rtl_int12.npz", iq=iq, hysteresis=np.int64(2), timing_offset=np.int64(0),
        **{k: np.int64(v) for k, v in kw.items()}, **_rtl_state_dict(st, det),
    )
    print("minn_rtl int12", _rtl_state_dict(st, det)["events"].tolist())


def _aa_pack(res) -> dict:
    ev_i = np.array(
        [[e.peak_index, e.gate_start, e.gate_end, e.frame_start] for e in res.events], dtype=np.int64
    ).reshape(-1, 4)
    ev_f = np.array(
        [[e.P_at_peak.real, e.P_at_peak.imag, e.M_at_peak, e.cfo_hz] for e in res.events], dtype=np.float64
    ).reshape(-1, 4)
    return dict(P=res.state.P, R=res.state.R, M=res.state.M, valid=res.state.valid, ev_i=ev_i, ev_f=ev_f)


def gen_sync_aa(out: Path, ref: str) -> None:
    import sync_aa as aa
    # (a) the docs scenario: [500 zeros][preamble][500 zeros], 1 antenna, no noise, L=512
    pre, _, _ = aa.build_aa_preamble(1024)
    tx = np.concatenate([np.zeros(500, complex), pre, np.zeros(500, complex)])
    docs = {}
    for name, cfo in (("clean", 0.0), ("cfo", 500.0)):
        rx = aa.apply_cfo(tx, cfo, aa.SAMPLE_RATE_HZ) if cfo else tx
        res = aa.aa_detect_streaming(rx, L=512)
        for k, v in _aa_pack(res).items():
            docs[f"{name}_{k}"] = v
        docs[f"{name}_rx"] = rx
        print("sync_aa docs", name, docs[f"{name}_ev_i"].tolist(), docs[f"{name}_ev_f"].tolist())
    # the reference's own golden vectors (docs/*.csv, docs/*.hex) stored as arrays
    d = Path(ref) / "docs"
    docs["csv_clean"] = np.genfromtxt(d / "detector_test_vector.csv", delimiter=",", comments="#", skip_header=4)
    docs["csv_cfo"] = np.genfromtxt(d / "detector_cfo_test_vector.csv", delimiter=",", comments="#", skip_header=3)
    docs["csv_preamble"] = np.genfromtxt(d / "preamble_test_vector.csv", delimiter=",", skip_header=1)
    hexw = [int(l.split()[0], 16) for l in (d / "preamble_test_vector.hex").read_text().splitlines()
            if l.strip() and not l.startswith("//")]
    docs["hex_preamble"] = np.asarray(hexw, dtype=np.int64)
    docs["preamble"] = pre
    docs["preamble_q12"] = aa.quantize_adc(pre, full_scale=2.0)
    np.savez_compressed(out / "sync_aa_docs.npz", **docs)

    # (b) grid cases through the real run_single_test (capture the detector's in/out)
    orig = aa.aa_detect_streaming
    cases = [(10.0, None, 2.0, 1024), (0.0, "cir1", 2.0, 1024), (10.0, "cir2", 1.0, 512),
             (5.0, None, 1.5, 256), (-5.0, "cir1", 2.0, 1024)]
    for i, (snr, ch, fs, plen) in enumerate(cases):
        box = {}

        def spy(rx, L=aa.PREAMBLE_HALF_LEN, **kw):
            r = orig(rx, L=L, **kw)
            box.update(rx=np.array(rx), L=np.int64(L), **_aa_pack(r))
            return r

        aa.aa_detect_streaming = spy
        try:
            tr = aa.run_single_test(snr, ch, fs, preamble_length=plen)
        finally:
            aa.aa_detect_streaming = orig
        np.savez_compressed(
            out / f"sync_aa_grid{i}.npz", snr=np.float64(snr), fs_ratio=np.float64(fs),
            timing_error=np.int64(tr.timing_error), detected=np.bool_(tr.detected),
            cfo_est=np.float64(tr.cfo_estimated_hz), **box,
        )
        print("sync_aa grid", i, snr, ch, fs, plen, tr.detected, tr.timing_error, tr.cfo_estimated_hz)


def gen_detector_cases(out: Path) -> None:
    """Function-level detector KATs on synthetic metrics (edge rules of SURVEY.md §8a)."""
    import sc, minn, combined_sc_min as c
    rng = np.random.default_rng(11)
    d = {}
    # plateau finder: three shapes exercising the three return paths (sc.py:106-114,117-133,136-146)
    n = 3000
    base = 0.02 * rng.random(n)
    m1 = base.copy(); m1[1000:1500] += 0.8 + 0.05 * rng.random(500); m1[1500:1600] += np.linspace(0.8, 0, 100)
    m2 = base.copy(); m2[500:2900] += 0.5 + 0.001 * rng.random(2400)          # no early drop inside cp
    m3 = np.zeros(n); m3[100] = 1.0                                             # isolated spike
    for k, m in (("p1", m1), ("p2", m2), ("p3", m3)):
        d[f"{k}_M"] = m
        d[f"{k}_end"] = np.int64(sc.find_plateau_end_from_metric(m, 512, lookahead=128, smooth_win=16))
        d[f"{k}_end_default"] = np.int64(sc.find_plateau_end_from_metric(m, 512))
    # minn peak finder incl. ties on run length and search bounds (minn.py:131-205)
    g1 = base.copy(); g1[400:420] += 1.0; g1[900:920] += 1.0; g1[1500:1510] += 2.0
    for k, (m, kw) in {
        "g1": (g1, dict(smooth_win=16, gate_threshold=0.5)),
        "g2": (g1, dict(smooth_win=1, gate_threshold=0.3, search_bounds=(800, 1200))),
        "g3": (g1, dict(smooth_win=8, gate_threshold=0.5, search_bounds=(2000, 100))),
        "g4": (g1, dict(smooth_win=4, gate_threshold=0.99)),
    }.items():
        pk, mask, ms = minn.find_minn_peak(m, **kw)
        d[f"{k}_M"], d[f"{k}_peak"], d[f"{k}_gate"], d[f"{k}_Ms"] = m, np.int64(pk), _segments(mask), ms
    # gated first-segment peak (combined_sc_min.py:183-259)
    gate = np.zeros(n, bool); gate[395:430] = True; gate[890:1000] = True
    d["c1_M"], d["c1_gate"] = g1, gate
    d["c1_peak"] = np.int64(c.find_minn_peak(g1, smooth_win=16, gate_mask=gate))
    d["c2_peak"] = np.int64(c.find_minn_peak(g1, smooth_win=16, gate_mask=gate, search_bounds=(850, 2000)))
    np.savez_compressed(out / "detector_cases.npz", **d)


def gen_cir_and_scale(out: Path) -> None:
    """CIR taps (channel_models/cir{1,2}.csv through channel.load_measured_cir) and a small BASELINE-cfg-2 style
    parity set made with the REAL channel.py / core.apply_cfo: tiled sc.py frames, cir1/cir2 ch1, AWGN, CFO,
    cast to complex64; expected outputs from the unmodified sc.py functions on the cast input."""
    import channel, core, sc
    np.savez_compressed(HERE.parent / "ofdm_sync_math_b200" / "data" / "channel_models.npz", cir1=channel.load_measured_cir("cir1"), cir2=channel.load_measured_cir("cir2"))
    n_samples, n_frames = 24576, 4
    xs, Ms, ends = [], [], []
    for f in range(n_frames):
        rng = np.random.default_rng(f)
        pre = sc.build_sc_preamble(rng, include_cp=True)
        pilot, _ = core.build_random_qpsk_symbol(rng, include_cp=True)
        data, _ = core.build_random_qpsk_symbol(rng, include_cp=True)
        fr = np.concatenate((np.zeros(core.TX_PRE_PAD_SAMPLES, complex), pre, pilot, data))
        cir = channel.load_measured_cir("cir1" if f % 2 == 0 else "cir2")[1:2]
        n_tx = n_samples - (cir.shape[1] - 1)
        tx = np.tile(fr, (n_tx + fr.size - 1) // fr.size)[:n_tx]
        rx = channel.apply_channel(tx, [0.0, 5.0, 10.0, 15.0][f], rng, channel_impulse_response=cir)
        rx = core.apply_cfo(rx, [-10e3, -3e3, 3e3, 10e3][f], core.SAMPLE_RATE_HZ)[0].astype(np.complex64)
        M, P, R = sc.sc_streaming_metric(rx.astype(np.complex128))
        xs.append(rx); Ms.append(M)
        ends.append(sc.find_plateau_end_from_metric(M, 512, lookahead=128, smooth_win=16))
    np.savez_compressed(out / "scale_sc.npz", x=np.stack(xs), M=np.stack(Ms).astype(np.float64), plateau_end=np.asarray(ends, np.int64))
    print("scale_sc", ends)


def gen_channel_and_cfo(out: Path) -> None:
    """the impairment chain channel.apply_channel (channel.py:51-98) -> core.apply_cfo (core.py:123-138)
    -> sync_aa.quantize_adc (sync_aa.py:263-291) and the CP-correlation CFO estimators (core.py:179-336), run UNMODIFIED on
    seeded inputs.  The AWGN is reproducible from `seed`: channel._compute_awgn_noise draws rng.standard_normal(shape) for
    the real part, then again for the imaginary part."""
    import channel, core, sc, sync_aa
    cases = {}
    for ci, (cir_name, snr, cfo, seed) in enumerate([("cir1", 10.0, 1000.0, 11), ("cir2", 0.0, -7500.0, 12), (None, 20.0, 250.0, 13)]):
        rng = np.random.default_rng(seed)
        pre = sc.build_sc_preamble(rng, include_cp=True)
        pilot, _ = core.build_random_qpsk_symbol(rng, include_cp=True)
        data, _ = core.build_random_qpsk_symbol(rng, include_cp=True)
        tx = np.concatenate((np.zeros(core.TX_PRE_PAD_SAMPLES, complex), pre, pilot, data, np.zeros(300, complex)))
        cir = None if cir_name is None else channel.load_measured_cir(cir_name)
        rng2 = np.random.default_rng(1000 + seed)
        rx = channel.apply_channel(tx, snr, rng2, channel_impulse_response=cir)
        rx_cfo = core.apply_cfo(rx, cfo, core.SAMPLE_RATE_HZ)
        fs_adc = 2.0 * float(np.sqrt(np.mean(np.abs(rx_cfo) ** 2)))
        rx_q = sync_aa.quantize_adc(rx_cfo, fs_adc, 12)
        pilot_cp_start = core.TX_PRE_PAD_SAMPLES + pre.size + (0 if cir is None else 170)
        est = {}
        for d_name, start in (("exact", pilot_cp_start), ("early", pilot_cp_start - 37), ("edge", 5), ("late", rx_cfo.shape[1] - 2560 - 120)):
            r = [core.estimate_cfo_from_cp(rx_cfo, start, core.N_FFT, core.CYCLIC_PREFIX, core.SAMPLE_RATE_HZ),
                 core.estimate_cfo_from_cp_robust(rx_cfo, start, core.N_FFT, core.CYCLIC_PREFIX, core.SAMPLE_RATE_HZ),
                 core.estimate_cfo_from_cp_peak(rx_cfo, start, core.N_FFT, core.CYCLIC_PREFIX, core.SAMPLE_RATE_HZ)]
            c4, d4 = core.estimate_cfo_from_cp_peak_with_index(rx_cfo, start, core.N_FFT, core.CYCLIC_PREFIX, core.SAMPLE_RATE_HZ, span=100)
            d5 = core.find_cp_start_via_corr(rx_cfo, start, core.N_FFT, core.CYCLIC_PREFIX, search_half=300)
            est[d_name] = np.array([start, *r, c4, d4, d5], dtype=np.float64)
        cases.update({f"tx{ci}": tx, f"cir{ci}": np.zeros((0, 0), complex) if cir is None else cir, f"snr{ci}": snr, f"cfo{ci}": cfo,
                      f"noise_seed{ci}": 1000 + seed, f"rx{ci}": rx, f"rx_cfo{ci}": rx_cfo, f"rx_q{ci}": rx_q, f"fs_adc{ci}": fs_adc,
                      **{f"est_{k}{ci}": v for k, v in est.items()}})
        print("channel/cfo", ci, cir_name, rx.shape, {k: v[:4].round(2).tolist() for k, v in est.items()})
    np.savez_compressed(out / "channel_cfo.npz", **cases)


def gen_rx_chain(out: Path) -> None:
    """the stage after the detector in every script -- CFO correction, pilot FFT -> LS channel estimate ->
    phase-slope timing -> equalise -> gain alignment -> EVM (core.py:171-176, 339-370, 443-469; call sites sc.py:286-309,
    minn.py:546-568).  Captured from the locals of the unmodified run_simulation() of sc.py (1 branch) and minn.py (2)."""
    import sc, minn
    for mod, name in ((sc, "sc"), (minn, "minn")):
        for ch, sub in SCEN:
            loc = _run_capture(mod, "run_simulation", (ch, sub))
            tag = ch or "awgn"
            np.savez_compressed(
                out / f"rxchain_{name}_{tag}.npz",
                rx=np.atleast_2d(loc["rx_samples"]), pilot_cp_start=np.int64(loc["pilot_cp_start"]), cfo_est_hz=np.float64(loc["cfo_est_hz"]),
                pilot_used=loc["pilot_used"], data_used=loc["data_used"], h_est=loc["h_est"], xhat_aligned=loc["xhat_aligned"],
                gain=np.complex128(loc["gain"]), evm_rms=np.float64(loc["evm_rms"]), evm_db=np.float64(loc["evm_db"]),
                slope=np.float64(loc["slope_rad_per_bin"]), sto=np.float64(loc["timing_offset_samples"]),
            )
            print("rxchain", name, tag, int(loc["pilot_cp_start"]), float(loc["evm_db"]), float(loc["timing_offset_samples"]))


def gen_aa_grid(out: Path) -> None:
    """the 135-case grid of sync_aa.main (sync_aa.py:1102-1109 -> run_grid_test :829-897 -> run_single_test
    :669-823), run UNMODIFIED and serially; one row of TestResult fields per case, in the reference's loop order."""
    import contextlib
    import io
    import time
    import sync_aa as aa
    kw = dict(snr_values=[-5, 0, 5, 10, 15], channels=[None, "cir1", "cir2"], full_scale_ratios=[0.5, 1.0, 2.0],
              preamble_lengths=list(aa.PREAMBLE_LENGTHS), cfo_hz=500.0, plot_samples=False)
    t0 = time.perf_counter()
    with contextlib.redirect_stdout(io.StringIO()):
        res = aa.run_grid_test(**kw)
    dt = time.perf_counter() - t0
    chan_code = {"awgn": 0, "cir1": 1, "cir2": 2}
    np.savez_compressed(
        out / "aa_grid.npz",
        snr_db=np.array([r.snr_db for r in res], dtype=np.float64), channel=np.array([chan_code[r.channel] for r in res], dtype=np.int64),
        fs_ratio=np.array([r.full_scale_ratio for r in res], dtype=np.float64),
        preamble_length=np.array([r.preamble_length for r in res], dtype=np.int64),
        timing_error=np.array([r.timing_error for r in res], dtype=np.int64),
        cfo_estimated_hz=np.array([r.cfo_estimated_hz for r in res], dtype=np.float64),
        cfo_error_hz=np.array([r.cfo_error_hz for r in res], dtype=np.float64),
        detected=np.array([r.detected for r in res], dtype=np.bool_), num_events=np.array([r.num_events for r in res], dtype=np.int64),
        clipping_pct=np.array([r.clipping_pct for r in res], dtype=np.float64),
        effective_bits=np.array([r.effective_bits for r in res], dtype=np.float64),
        metric_peak=np.array([r.metric_peak for r in res], dtype=np.float64), reference_seconds=np.float64(dt),
    )
    print(f"aa grid: {len(res)} cases, {sum(r.detected for r in res)} detected, reference run_grid_test took {dt:.1f} s")


def gen_wire(out: Path, ref: str) -> None:
    """The RTL testbench's AXIS word packer (ref/test_minn_preamble_detector.py:41-47), executed UNMODIFIED: the module imports
    cocotb (absent here), so the one function and the INPUT_WIDTH constant it uses are pulled out of the parsed file."""
    import ast
    src = (Path(ref) / "ref" / "test_minn_preamble_detector.py").read_text()
    tree = ast.parse(src)
    keep = [n for n in tree.body
            if (isinstance(n, ast.FunctionDef) and n.name == "_pack_axis_samples")
            or (isinstance(n, ast.Assign) and any(isinstance(t, ast.Name) and t.id == "INPUT_WIDTH" for t in n.targets))]
    ns: dict = {}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "test_minn_preamble_detector.py", "exec"), ns)
    rng = np.random.default_rng(5)
    iq = rng.integers(-2048, 2048, size=(2, 4096, 2)).astype(np.int16)
    iq[:, :4] = [[[-2048, 2047], [2047, -2048], [0, -1], [-1, 0]]] * 2
    words = np.array([ns["_pack_axis_samples"](int(iq[0, i, 0]), int(iq[0, i, 1]), int(iq[1, i, 0]), int(iq[1, i, 1]))
                      for i in range(iq.shape[1])], dtype=np.int64)
    np.savez_compressed(out / "wire_axis.npz", iq=iq, words=words, input_width=np.int64(ns["INPUT_WIDTH"]))
    print("wire axis", iq.shape, hex(int(words[0])), "INPUT_WIDTH", ns["INPUT_WIDTH"])


def gen_sweeps(out: Path) -> None:
    """minn.compare_block_lengths (minn.py:754-871) at SNR 0 / 10 dB and minn_rtl.compare_q_values
    (minn_rtl.py:1493-1592), run UNMODIFIED for flat AWGN and cir1; scalar results per sweep point, plus the TX builders'
    outputs the host-side tests pin (minn_rtl.build_minn_preamble_generic for every sequence type)."""
    import time
    import minn
    import minn_rtl
    d = {}
    Ns, Qs = [256, 512, 1024, 2048], [64, 128, 256, 512]
    keys = ("peak", "par", "pmr", "timing_error", "preamble_len", "overhead_pct")
    t0 = time.perf_counter()
    for ch, _ in SCEN:
        tag = ch or "awgn"
        for snr in (0.0, 10.0):
            r = minn.compare_block_lengths(Ns, ch, snr)
            for k in keys:
                d[f"block_{tag}_snr{int(snr)}_{k}"] = np.array([r[N][k] for N in Ns], dtype=np.float64)
            print("block lengths", tag, snr, [r[N]["timing_error"] for N in Ns], [round(r[N]["peak"], 4) for N in Ns])
        d[f"block_{tag}_metric256"] = r[256]["metric"]
        r = minn_rtl.compare_q_values(Qs, ch)
        for k in keys:
            d[f"q_{tag}_{k}"] = np.array([r[Q][k] for Q in Qs], dtype=np.float64)
        print("q values", tag, [r[Q]["timing_error"] for Q in Qs], [round(r[Q]["peak"], 2) for Q in Qs])
    d["reference_seconds"] = np.float64(time.perf_counter() - t0)
    d["block_lengths"], d["q_values"] = np.array(Ns), np.array(Qs)
    types = ["bpsk_freq", "qpsk_freq", "zc_time", "zc_freq", "chirp", "gold", "const", "random_phase"]
    d["seq_types"] = np.array(types)
    for t in types:
        d[f"pre_{t}"] = minn_rtl.build_minn_preamble_generic(t, np.random.default_rng(3), Q=64)
    d["pre_param_256"] = minn.build_minn_preamble_parameterized(np.random.default_rng(0), 256, 64)
    np.savez_compressed(out / "sweeps.npz", **d)
    print(f"sweeps: reference took {float(d['reference_seconds']):.1f} s")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=str(HERE.parent / "tests" / "golden"))
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    out = Path(a.out).resolve()
    out.mkdir(parents=True, exist_ok=True)
    _setup(a.ref)
    gens = dict(sc=gen_sc, minn=gen_minn, park=gen_park, combined=gen_combined, zc=gen_zc, zc_v2=gen_zc_v2,
                zc_freq=gen_zc_freq, minn_rtl=gen_minn_rtl, detector_cases=gen_detector_cases,
                cir_scale=gen_cir_and_scale, channel_cfo=gen_channel_and_cfo, rx_chain=gen_rx_chain,
                aa_grid=gen_aa_grid, sweeps=gen_sweeps)
    for name, fn in gens.items():
        if a.only and name not in a.only.split(","):
            continue
        fn(out)
    if not a.only or "sync_aa" in a.only.split(","):
        gen_sync_aa(out, a.ref)
    if not a.only or "wire" in a.only.split(","):
        gen_wire(out, a.ref)
    total = sum(p.stat().st_size for p in out.glob("*.npz"))
    print(f"wrote {len(list(out.glob('*.npz')))} fixtures, {total / 1e6:.1f} MB -> {out}")


if __name__ == "__main__":
    main()
