"""Pins the CPU oracle (oracle/) against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/gen_golden.py) and against the reference's own docs/*_test_vector files.
Float tolerance: 1e-11 relative to the array scale (numpy's internal summation order differs
from the oracle's sequential sums); indices, masks and event lists must be equal."""
import numpy as np
import pytest

from oracle import oracle as orc
from conftest import segments

RTOL = 1e-11


def close(a, b, rtol=RTOL):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    err = float(np.max(np.abs(a - b))) if b.size else 0.0
    assert err <= rtol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_sc(golden, tag):
    g = golden(f"sc_{tag}")
    M, P, R = orc.sc_streaming_metric(g["rx"])
    close(M, g["M"]); close(P, g["P"]); close(R, g["R"])
    for Min in (M, g["M"]):
        end = orc.find_plateau_end_from_metric(Min, int(g["cp_len"]), lookahead=int(g["lookahead"]),
                                               smooth_win=int(g["smooth_win"]))
        assert end == int(g["plateau_end"])
        assert max(end - int(g["sc_delta"]), 0) == int(g["coarse_start"])


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_minn(golden, tag):
    g = golden(f"minn_{tag}")
    M, P, R = orc.minn_streaming_metric(g["rx"])
    close(M, g["M"]); close(P, g["P"]); close(R, g["R"])
    pk, gate, Ms = orc.find_minn_peak(g["M"], smooth_win=int(g["smooth_win"]), gate_threshold=float(g["gate_threshold"]))
    assert pk == int(g["peak"]); assert np.array_equal(segments(gate), g["gate"]); close(Ms, g["Ms"], 1e-13)
    pk2, gate2, _ = orc.find_minn_peak(M, smooth_win=int(g["smooth_win"]), gate_threshold=float(g["gate_threshold"]))
    assert pk2 == int(g["peak"]); assert np.array_equal(segments(gate2), g["gate"])


def test_minn_param(golden):
    g = golden("minn_param")
    for n in (64, 100, 256, 1024):
        M, P, R = orc.minn_streaming_metric_parameterized(g["rx"], n)
        close(M, g[f"M_{n}"]); close(P, g[f"P_{n}"]); close(R, g[f"R_{n}"])
    M, P, R = orc.minn_streaming_metric_parameterized(g["rx"], 4096)   # L < N -> empty (minn.py:723-724)
    assert M.size == 0 and P.size == 0 and R.size == 0 and P.dtype == complex


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_park(golden, tag):
    g = golden(f"park_{tag}")
    ds, M, P, E = orc.park_streaming_metric(g["rx"])
    assert np.array_equal(ds, g["ds"]); close(M, g["M"]); close(P, g["P"]); close(E, g["E"])
    c = int(ds[int(np.argmax(M))])
    assert c == int(g["det_center"]) and max(c - 1024, 0) == int(g["det_symbol_start"])


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_combined(golden, tag):
    g = golden(f"combined_{tag}")
    M, P, R = orc.minn_streaming_metric(g["rx"])
    Msc, Psc, Rsc = orc.schmidl_cox_streaming_metric(g["rx"])
    close(M, g["M"]); close(Msc, g["M_sc"]); close(Psc, g["P_sc"]); close(Rsc, g["R_sc"])
    gate = orc.sc_gate(Msc, float(g["sc_gate_threshold"]))
    assert np.array_equal(segments(gate), g["sc_gate"])
    idx = np.flatnonzero(gate)
    assert (idx[0], idx[-1] + 1) == tuple(g["sc_gate_span"])
    assert orc.find_minn_peak_gated(M, smooth_win=int(g["smooth_win"]), gate_mask=gate) == int(g["peak"])


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_zc(golden, tag):
    g = golden(f"zc_{tag}")
    corr, peak, start = orc.zc_correlation(g["rx"], g["ref"])
    close(corr, g["corr"]); assert peak == int(g["peak"]) and start == int(g["start"])


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_zc_v2(golden, tag):
    g = golden(f"zc_v2_{tag}")
    mag = orc.zc_v2_corr_mag(g["rx"], g["ref"])
    close(mag, g["corr_mag"])
    for m in (mag, g["corr_mag"]):
        st = orc.zc_streaming_detection(m, int(g["window_size"]), int(g["thresh_value"]),
                                        int(g["thresh_frac_bits"]), float(g["min_corr_mag"]))
        close(st.local_sum, g["local_sum"], 1e-10)
        assert np.array_equal(st.metric_valid, g["valid"]); assert np.array_equal(st.above_threshold, g["above"])
        ev, vals, gm = orc.detect_zc_peaks(st, g["ref"].size, int(g["hysteresis"]))
        assert np.array_equal(ev, g["events"]); close(vals, g["event_values"]); assert np.array_equal(segments(gm), g["gate"])


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_zc_freq(golden, tag):
    g = golden(f"zc_freq_{tag}")
    n = 600   # the direct-DFT restatement is O(62*2048) per offset: check a window around the peak + the start
    pk = int(g["peak"])
    lo = max(pk - n // 2, 0)
    rx = g["rx"][:, lo: lo + n + 2559]
    m = orc.compute_frequency_metric(rx, g["bin_indices"], g["template"], float(g["template_energy"]))
    close(m, g["metric"][lo: lo + m.size], 1e-10)
    assert lo + int(np.argmax(m)) == pk
    with pytest.raises(ValueError):
        orc.compute_frequency_metric(g["rx"][:, :2559], g["bin_indices"], g["template"], float(g["template_energy"]))


@pytest.mark.parametrize("tag", ["cir1", "awgn", "int12"])
def test_minn_rtl(golden, tag):
    g = golden(f"minn_rtl_{tag}")
    if tag == "int12":
        rx = g["iq"][..., 0].astype(np.float64) + 1j * g["iq"][..., 1].astype(np.float64)
    else:
        rx = g["rx"]
    kw = dict(smooth_shift=int(g["smooth_shift"]), threshold_value=int(g["threshold_value"]),
              threshold_frac_bits=int(g["threshold_frac_bits"]), quarter_len=int(g["quarter_len"]))
    st = orc.minn_rtl_streaming_metric(rx, **kw)
    exact = tag == "int12"       # integer-valued input: every float64 op is exact -> bit-equal
    for k in ("corr_total", "corr_positive", "smooth_metric", "energy_total", "corr_scaled", "energy_scaled"):
        if exact:
            assert np.array_equal(st[k], g[k]), k
        else:
            close(st[k], g[k], 1e-12)
    assert np.array_equal(st["metric_valid"], g["metric_valid"]); assert np.array_equal(st["above"], g["above"])
    ev, segs = orc.detect_minn_rtl(st, hysteresis=int(g["hysteresis"]), timing_offset=int(g["timing_offset"]))
    assert np.array_equal(ev, g["events"]); assert np.array_equal(segs, g["gate_segments"])


def test_minn_rtl_int_model_vs_float_mirror(golden):
    """Integer SV model == minn_rtl.py on int12 input for corr/energy/valid (exact); the floor-shift
    smoother differs from the float smoother by design: bounded, flags compared."""
    g = golden("minn_rtl_int12")
    kw = dict(smooth_shift=3, threshold_value=3276, threshold_frac_bits=15, quarter_len=512)
    d = orc.minn_rtl_int(g["iq"], **kw)
    assert np.array_equal(d["corr_total"], g["corr_total"].astype(np.int64))
    assert np.array_equal(d["energy_total"], g["energy_total"].astype(np.int64))
    assert np.array_equal(d["corr_positive"], g["corr_positive"].astype(np.int64))
    assert np.array_equal(d["metric_valid"], g["metric_valid"])
    assert np.max(np.abs(d["smooth_metric"] - g["smooth_metric"])) < 8.0   # < 2^shift LSB
    ev, segs = orc.detect_minn_rtl(d, hysteresis=2, timing_offset=0)
    assert abs(int(ev[0, 0]) - int(g["events"][0, 0])) <= 16              # the testbench's own criterion


def test_sync_aa_docs_vectors(golden):
    """docs/detector_test_vector.csv + docs/detector_cfo_test_vector.csv + docs/preamble_test_vector.{csv,hex}
    (these pin sync_aa.aa_detect_streaming)."""
    g = golden("sync_aa_docs")
    for name, csv in (("clean", g["csv_clean"]), ("cfo", g["csv_cfo"])):
        r = orc.aa_detect_streaming(g[f"{name}_rx"], L=512)
        # P: bit-exact against the reference run (pure scalar recurrence).  R, M: numpy's |z|^2 goes
        # through its own hypot and is not bit-reproducible (see ofs_oracle.c sq_abs) -> few ulp.
        assert np.array_equal(r["P"], g[f"{name}_P"]); close(r["R"], g[f"{name}_R"], 4e-15)
        assert np.max(np.abs(r["M"] - g[f"{name}_M"])) <= 4e-15; assert np.array_equal(r["valid"], g[f"{name}_valid"])
        assert np.array_equal(r["ev_i"], g[f"{name}_ev_i"]); close(r["ev_f"], g[f"{name}_ev_f"], 4e-15)
        assert r["ev_i"].tolist() == [[1523, 1210, 2024, 500]]
        # the CSV rows (samples 1000..1599), to their printed precision
        s = csv[:, 0].astype(int)
        assert np.max(np.abs(r["M"][s] - csv[:, 1])) <= 5.1e-9
        assert np.max(np.abs(r["P"][s].real - csv[:, 2])) <= 5.1e-3
        assert np.max(np.abs(r["P"][s].imag - csv[:, 3])) <= 5.1e-3
        assert np.max(np.abs(np.abs(r["P"][s]) ** 2 - csv[:, 4])) <= 5.1e-3 + 1e-9 * np.max(csv[:, 4])
        if name == "clean":
            assert np.max(np.abs(r["R"][s] - csv[:, 5])) <= 5.1e-3
        else:
            assert np.max(np.abs(np.angle(r["P"][s]) - csv[:, 5])) <= 5.1e-9
    assert abs(g["cfo_ev_f"][0, 3] - 500.0) < 1e-9
    # preamble vectors: float columns, int12 columns == quantize_adc(x, 2.0) * 1024, hex packing
    pc = g["csv_preamble"]
    assert np.max(np.abs(g["preamble"].real - pc[:, 1])) < 6e-11 and np.max(np.abs(g["preamble"].imag - pc[:, 2])) < 6e-11
    q = np.round(g["preamble_q12"] * 1024.0)
    assert np.array_equal(q.real.astype(int), pc[:, 3].astype(int)); assert np.array_equal(q.imag.astype(int), pc[:, 4].astype(int))
    words = ((pc[:, 3].astype(np.int64) & 0xFFF) << 12) | (pc[:, 4].astype(np.int64) & 0xFFF)
    assert np.array_equal(words, g["hex_preamble"])


@pytest.mark.parametrize("i", range(5))
def test_sync_aa_grid(golden, i):
    g = golden(f"sync_aa_grid{i}")
    r = orc.aa_detect_streaming(g["rx"], L=int(g["L"]))
    for k in ("P", "valid", "ev_i"):
        assert np.array_equal(r[k], g[k]), k
    for k in ("R", "M", "ev_f"):
        close(r[k], g[k], 4e-15)
    assert bool(g["detected"]) == (r["ev_i"].shape[0] > 0)


def test_detector_cases(golden):
    g = golden("detector_cases")
    for k in ("p1", "p2", "p3"):
        assert orc.find_plateau_end_from_metric(g[f"{k}_M"], 512, lookahead=128, smooth_win=16) == int(g[f"{k}_end"])
        assert orc.find_plateau_end_from_metric(g[f"{k}_M"], 512) == int(g[f"{k}_end_default"])
    kws = dict(g1=dict(smooth_win=16, gate_threshold=0.5), g2=dict(smooth_win=1, gate_threshold=0.3, search_bounds=(800, 1200)),
               g3=dict(smooth_win=8, gate_threshold=0.5, search_bounds=(2000, 100)), g4=dict(smooth_win=4, gate_threshold=0.99))
    for k, kw in kws.items():
        pk, gate, Ms = orc.find_minn_peak(g[f"{k}_M"], **kw)
        assert pk == int(g[f"{k}_peak"]), k
        assert np.array_equal(segments(gate), g[f"{k}_gate"]); close(Ms, g[f"{k}_Ms"], 1e-13)
    assert orc.find_minn_peak_gated(g["c1_M"], smooth_win=16, gate_mask=g["c1_gate"]) == int(g["c1_peak"])
    assert orc.find_minn_peak_gated(g["c1_M"], smooth_win=16, gate_mask=g["c1_gate"], search_bounds=(850, 2000)) == int(g["c2_peak"])


def test_edge_cases():
    z = np.zeros(100, complex)
    for f in (orc.sc_streaming_metric, orc.minn_streaming_metric, orc.schmidl_cox_streaming_metric):
        M, P, R = f(z)
        assert M.size == 0 and P.size == 0 and R.size == 0
    ds, M, P, E = orc.park_streaming_metric(np.zeros(2048, complex))
    assert ds.size == 0 and M.size == 0
    assert orc.find_plateau_end_from_metric(np.zeros(0), 512) == 0
    with pytest.raises(ValueError):
        orc.find_minn_peak(np.zeros(0))
    with pytest.raises(ValueError):
        orc.find_minn_peak(np.zeros(10))
    assert orc.find_minn_peak_gated(np.zeros(0)) == 0
    with pytest.raises(ValueError):
        orc.find_minn_peak_gated(np.ones(10))
    with pytest.raises(ValueError):
        orc.find_minn_peak_gated(np.ones(10), gate_mask=np.zeros(10, bool))


def test_prefix_scale_path_matches_literal():
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(6000) + 1j * rng.standard_normal(6000)).astype(np.complex64)
    for kind, fn in ((0, orc.sc_streaming_metric), (1, orc.schmidl_cox_streaming_metric), (2, orc.minn_streaming_metric)):
        M, P, R = orc.metric_prefix_c64(x, 2048, kind, want_pr=True)
        M0, P0, R0 = fn(x.astype(np.complex128))
        close(P, P0, 1e-12 * 50); close(R, R0, 1e-12)
        assert np.max(np.abs(M - M0) / np.maximum(M0, 1e-6)) < 1e-9


# ------------------------------------------------------------------------------------------- SURVEY 8(f) rows 1-2
@pytest.mark.parametrize("ci", [0, 1, 2])
def test_oracle_channel_chain_vs_reference(golden, ci):
    """channel.apply_channel -> core.apply_cfo -> sync_aa.quantize_adc and the CP-CFO estimators: oracle restatement vs the
    outputs of the unmodified reference (oracle/gen_golden.py::gen_channel_and_cfo)."""
    g = golden("channel_cfo")
    tx, cir = g[f"tx{ci}"], g[f"cir{ci}"]
    cir = None if cir.size == 0 else cir
    rx_ref = g[f"rx{ci}"]
    unit = orc.unit_noise_like_reference(int(g[f"noise_seed{ci}"]), rx_ref.shape)
    rx = orc.channel_apply(tx, cir, float(g[f"snr{ci}"]), unit)
    assert rx.shape == rx_ref.shape and np.abs(rx - rx_ref).max() <= 1e-12 * np.abs(rx_ref).max()
    rc = orc.apply_cfo(rx, float(g[f"cfo{ci}"]), 30.72e6)
    assert np.abs(rc - g[f"rx_cfo{ci}"]).max() <= 1e-12 * np.abs(rx_ref).max()
    q, codes = orc.quantize_adc(g[f"rx_cfo{ci}"], float(g[f"fs_adc{ci}"]), 12)
    assert np.array_equal(q, g[f"rx_q{ci}"])
    assert codes.min() >= -2048 and codes.max() <= 2047
    for name in ("exact", "early", "edge", "late"):
        e = g[f"est_{name}{ci}"]
        start = int(e[0])
        x = g[f"rx_cfo{ci}"]
        assert abs(orc.estimate_cfo_from_cp(x, start, 2048, 512, 30.72e6) - e[1]) <= 1e-6
        assert abs(orc.estimate_cfo_from_cp_robust(x, start, 2048, 512, 30.72e6) - e[2]) <= 1e-6
        c, d = orc.estimate_cfo_from_cp_peak_with_index(x, start, 2048, 512, 30.72e6)
        assert abs(c - e[3]) <= 1e-6
        c, d = orc.estimate_cfo_from_cp_peak_with_index(x, start, 2048, 512, 30.72e6, span=100)
        assert abs(c - e[4]) <= 1e-6 and d == int(e[5])
        assert orc.estimate_cfo_from_cp_peak_with_index(x, start, 2048, 512, 1.0, span=300)[1] == int(e[6])


@pytest.mark.parametrize("name", ["sc_cir1", "sc_awgn", "minn_cir1", "minn_awgn"])
def test_oracle_rx_chain_vs_reference(golden, name):
    """CFO correction -> pilot LS estimate -> phase-slope timing -> equalise -> align -> EVM, oracle vs the locals
    of the unmodified run_simulation() (oracle/gen_golden.py::gen_rx_chain)."""
    g = golden(f"rxchain_{name}")
    r = orc.rx_chain(g["rx"], int(g["pilot_cp_start"]), float(g["cfo_est_hz"]), g["pilot_used"], g["data_used"])
    assert np.abs(r["h_est"] - g["h_est"]).max() <= 1e-10 * np.abs(g["h_est"]).max()
    assert np.abs(r["xhat"] - g["xhat_aligned"]).max() <= 1e-9 * np.abs(g["xhat_aligned"]).max()
    assert abs(r["gain"] - complex(g["gain"])) <= 1e-10 * abs(complex(g["gain"]))
    assert abs(r["evm_rms"] - float(g["evm_rms"])) <= 1e-10 and abs(r["evm_db"] - float(g["evm_db"])) <= 1e-8
    assert abs(r["slope"] - float(g["slope"])) <= 1e-10 and abs(r["sto"] - float(g["sto"])) <= 1e-7


def test_oracle_wire_formats_vs_reference_vectors(golden):
    """12-bit wire words: docs/preamble_test_vector.hex (with the csv's integer columns) and the words of
    the RTL testbench's own packer (ref/test_minn_preamble_detector.py:41-47, tests/golden/wire_axis.npz)."""
    d = golden("sync_aa_docs")
    _, codes = orc.quantize_adc(d["preamble"], 2.0)
    assert np.array_equal(codes, d["csv_preamble"][:, 3:5].astype(np.int16))
    assert np.array_equal(orc.wire_pack_hex24(codes), d["hex_preamble"])
    assert np.array_equal(orc.wire_unpack_hex24(d["hex_preamble"]), codes)
    g = golden("wire_axis")
    assert int(g["input_width"]) == 12
    assert np.array_equal(orc.wire_pack_axis48(g["iq"]), g["words"])
    assert np.array_equal(orc.wire_unpack_axis48(g["words"]), g["iq"])


def test_aa_grid_fixture_shape(golden):
    """The 135-case grid fixture (sync_aa.main's parameters, sync_aa.py:1102-1109) in the reference's loop order."""
    g = golden("aa_grid")
    assert g["detected"].shape == (135,) and g["preamble_length"][0] == 1024 and g["preamble_length"][-1] == 256
    assert np.all(g["num_events"][~g["detected"]] == 0) and np.all(g["num_events"][g["detected"]] >= 1)
    assert np.all(g["timing_error"][~g["detected"]] == 0)


def test_sweep_tx_builders_vs_reference(golden):
    """Host-side transmit builders of the sweep layer against the reference's outputs for the same seeds:
    minn_rtl.build_minn_preamble_generic (minn_rtl.py:335-358, all sequence types), minn.build_minn_preamble_parameterized
    (minn.py:656-694), sync_aa.build_aa_preamble (sync_aa.py:160-235)."""
    from ofdm_sync_math_b200 import minn, minn_rtl, sync_aa
    g = golden("sweeps")
    for t in g["seq_types"].tolist():
        got = minn_rtl.build_minn_preamble_generic(t, np.random.default_rng(3), Q=64)
        assert got.shape == (320,) and np.abs(got - g[f"pre_{t}"]).max() <= 1e-14, t
    assert np.array_equal(minn.build_minn_preamble_parameterized(np.random.default_rng(0), 256, 64), g["pre_param_256"])
    with pytest.raises(ValueError):
        minn.build_minn_preamble_parameterized(np.random.default_rng(0), 250, 64)
    with pytest.raises(ValueError):
        minn_rtl.build_minn_preamble_generic("nope", np.random.default_rng(0))
    with pytest.raises(ValueError):
        minn_rtl.build_minn_preamble_generic("qpsk_freq", None)
    d = golden("sync_aa_docs")
    pre, zc, papr = sync_aa.build_aa_preamble(1024)
    assert np.abs(pre - d["preamble"]).max() <= 1e-14 and zc.size == 300
    for total in (512, 256):
        p2, _, _ = sync_aa.build_aa_preamble(total)
        assert p2.size == total and np.abs(p2[: total // 2] - p2[total // 2:]).max() <= 1e-12
    with pytest.raises(ValueError):
        sync_aa.build_aa_preamble(300)


def test_oracle_aa_grid_vs_reference_run(golden):
    """The whole 135-case grid of sync_aa.main through the oracle (C detector + numpy impairment chain) against the unmodified
    reference run (tests/golden/aa_grid.npz): detection, event count and timing error equal, CFO 1e-6 Hz, peak metric 1e-9,
    clipping statistics 1e-9.  The transmit side comes from the host builders, pinned separately against the reference."""
    from ofdm_sync_math_b200 import sync_aa, synth
    g = golden("aa_grid")
    cirs = synth.load_cirs()
    cache = {}
    for i in range(g["detected"].size):
        plen, ch = int(g["preamble_length"][i]), int(g["channel"][i])
        if (plen, ch) not in cache:
            rng = np.random.default_rng(42)
            pre, _, _ = sync_aa.build_aa_preamble(plen)
            pilot, _ = sync_aa.build_random_qpsk_symbol(rng)
            data, _ = sync_aa.build_random_qpsk_symbol(rng)
            tx = np.concatenate((np.zeros(500, complex), pre, pilot, data, np.zeros(500, complex)))
            cir = np.ones((2, 1), complex) if ch == 0 else cirs["cir1" if ch == 1 else "cir2"][:2]
            n_out = tx.size + cir.shape[1] - 1
            unit = np.stack([rng.standard_normal(n_out) + 1j * rng.standard_normal(n_out) for _ in range(2)])
            off = 0 if ch == 0 else int(np.argmax(np.sum(np.abs(cir) ** 2, axis=0)))
            cache[(plen, ch)] = (tx, cir, unit, off)
        tx, cir, unit, off = cache[(plen, ch)]
        r = orc.aa_single_test(tx, cir, float(g["snr_db"][i]), float(g["fs_ratio"][i]), plen // 2, unit, 500.0, true_start=500 + off)
        tag = f"case {i}"
        assert r["detected"] == bool(g["detected"][i]) and r["num_events"] == int(g["num_events"][i]), tag
        assert r["timing_error"] == int(g["timing_error"][i]), tag
        assert abs(r["cfo_estimated_hz"] - g["cfo_estimated_hz"][i]) <= 1e-6, tag
        assert abs(r["metric_peak"] - g["metric_peak"][i]) <= 1e-9 * max(abs(g["metric_peak"][i]), 1e-12), tag
        assert abs(r["clipping_pct"] - g["clipping_pct"][i]) <= 1e-9 and abs(r["effective_bits"] - g["effective_bits"][i]) <= 1e-9, tag


@pytest.mark.parametrize("ch", [None, "cir1"])
def test_oracle_block_length_sweep_vs_reference_run(golden, ch):
    """minn.compare_block_lengths (minn.py:754-871) through the oracle against the unmodified reference run
    (tests/golden/sweeps.npz): timing errors equal, peak / PAR / PMR 1e-9 relative."""
    from ofdm_sync_math_b200 import minn, sweeps
    g = golden("sweeps")
    tag = ch or "awgn"
    cir, delay = sweeps.channel_bank(ch)
    for snr in (0.0, 10.0):
        got = []
        for N in g["block_lengths"].tolist():
            rng = np.random.default_rng(0)
            pre = minn.build_minn_preamble_parameterized(rng, N, N // 4, include_cp=True)
            tx, frame_len = sweeps.two_frame_stream(pre, rng)
            B = 1 if cir is None else cir.shape[0]
            n_out = tx.size + (0 if cir is None else cir.shape[1] - 1)
            unit = orc.unit_noise_like_reference(0, (B, n_out))
            got.append(orc.block_length_point(tx, frame_len, cir, snr, N, unit, delay=delay))
        got = np.array(got)
        pre_ = f"block_{tag}_snr{int(snr)}"
        assert got[:, 3].astype(int).tolist() == g[f"{pre_}_timing_error"].astype(int).tolist()
        for col, key in enumerate(("peak", "par", "pmr")):
            ref = g[f"{pre_}_{key}"]
            assert np.all(np.abs(got[:, col] - ref) <= 1e-9 * np.abs(ref)), (pre_, key, got[:, col], ref)


@pytest.mark.parametrize("ch", [None, "cir1"])
def test_oracle_q_sweep_vs_reference_run(golden, ch):
    """minn_rtl.compare_q_values (minn_rtl.py:1493-1592) through the oracle against the unmodified reference run."""
    from ofdm_sync_math_b200 import minn_rtl, sweeps
    g = golden("sweeps")
    tag = ch or "awgn"
    cir, delay = sweeps.channel_bank(ch)
    got = []
    for Q in g["q_values"].tolist():
        rng = np.random.default_rng(0)
        pre = minn_rtl.build_minn_preamble_generic("qpsk_freq", rng, Q=Q)
        tx, _ = sweeps.two_frame_stream(pre, rng)
        B = 1 if cir is None else cir.shape[0]
        n_out = tx.size + (0 if cir is None else cir.shape[1] - 1)
        got.append(orc.q_value_point(tx, cir, 0.0, Q, orc.unit_noise_like_reference(0, (B, n_out)), delay=delay))
    got = np.array(got)
    assert got[:, 3].astype(int).tolist() == g[f"q_{tag}_timing_error"].astype(int).tolist()
    for col, key in enumerate(("peak", "par", "pmr")):
        ref = g[f"q_{tag}_{key}"]
        assert np.all(np.abs(got[:, col] - ref) <= 1e-9 * np.abs(ref)), (key, got[:, col], ref)


def test_sweep_peak_statistics_host_logic_vs_oracle():
    """sweeps.peak_statistics (torch, device-agnostic) against the oracle's restatement of minn.py:841-858 on CPU tensors:
    random metrics, peaks at the row ends, a guard that swallows the whole row (-> inf), all-zero noise floor (-> inf)."""
    import torch
    from ofdm_sync_math_b200 import sweeps
    rng = np.random.default_rng(0)
    rows = []
    for n, pk in ((5000, 2500), (5000, 0), (5000, 4999), (1400, 1350), (900, 450), (3000, 1500)):
        m = rng.random(n)
        if n == 3000:
            m[:] = 0.0; m[pk] = 1.0                       # zero noise floor
        rows.append((m, pk))
    for m, pk in rows:
        peak, par, pmr = sweeps.peak_statistics(torch.as_tensor(m)[None], torch.as_tensor([pk]))
        ref = orc.peak_noise_statistics(m, pk)
        got = (float(peak[0]), float(par[0]), float(pmr[0]))
        for a, b in zip(got, ref):
            assert (np.isinf(a) and np.isinf(b)) or abs(a - b) <= 1e-12 * abs(b), (m.size, pk, got, ref)
