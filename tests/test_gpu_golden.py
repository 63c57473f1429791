"""GPU parity, part 1: the drop-in modules (CUDA kernels behind the C ABI) against outputs of the unmodified
reference (tests/golden/*.npz).  Same assertions as tests/test_oracle_golden.py: float arrays to 1e-11 of the
array scale (float64 kernels; numpy's summation order differs), indices / masks / event lists equal."""
import numpy as np
import pytest

from conftest import segments

pytestmark = pytest.mark.gpu
RTOL = 1e-11


def close(a, b, rtol=RTOL):
    a = np.asarray(a); b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(float(np.max(np.abs(b))) if b.size else 0.0, 1e-300)
    err = float(np.max(np.abs(a - b))) if b.size else 0.0
    assert err <= rtol * scale, f"max err {err:.3e} vs scale {scale:.3e}"


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_sc(golden, tag):
    from ofdm_sync_math_b200 import sc
    g = golden(f"sc_{tag}")
    M, P, R = sc.sc_streaming_metric(g["rx"])
    assert M.dtype == np.float64 and P.dtype == np.complex128
    close(M, g["M"]); close(P, g["P"]); close(R, g["R"])
    for Min in (M, g["M"]):
        end = sc.find_plateau_end_from_metric(Min, int(g["cp_len"]), lookahead=int(g["lookahead"]), smooth_win=int(g["smooth_win"]))
        assert end == int(g["plateau_end"])
    assert sc.find_plateau_end_from_metric(np.zeros(0), 512) == 0
    M0, P0, R0 = sc.sc_streaming_metric(g["rx"][:, :100])
    assert M0.size == 0 and P0.size == 0 and R0.size == 0


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_minn(golden, tag):
    from ofdm_sync_math_b200 import minn
    g = golden(f"minn_{tag}")
    M, P, R = minn.minn_streaming_metric(g["rx"])
    close(M, g["M"]); close(P, g["P"]); close(R, g["R"])
    for Min in (M, g["M"]):
        pk, gate, Ms = minn.find_minn_peak(Min, smooth_win=int(g["smooth_win"]), gate_threshold=float(g["gate_threshold"]))
        assert pk == int(g["peak"]); assert np.array_equal(segments(gate), g["gate"]); close(Ms, g["Ms"], 1e-12)
    close(minn._trailing_average(g["M"], 16), g["Ms"], 1e-12)
    with pytest.raises(ValueError):
        minn.find_minn_peak(np.zeros(0))
    with pytest.raises(ValueError):
        minn.find_minn_peak(np.zeros(10))


def test_minn_param(golden):
    from ofdm_sync_math_b200 import minn
    g = golden("minn_param")
    for n in (64, 100, 256, 1024):
        M, P, R = minn.minn_streaming_metric_parameterized(g["rx"], n)
        close(M, g[f"M_{n}"]); close(P, g[f"P_{n}"]); close(R, g[f"R_{n}"])
    M, P, R = minn.minn_streaming_metric_parameterized(g["rx"], 4096)
    assert M.size == 0 and P.size == 0 and R.size == 0


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_park(golden, tag):
    from ofdm_sync_math_b200 import park
    g = golden(f"park_{tag}")
    ds, M, P, E = park.park_streaming_metric(g["rx"])
    assert np.array_equal(ds, g["ds"]); close(M, g["M"]); close(P, g["P"]); close(E, g["E"])
    assert int(ds[int(np.argmax(M))]) == int(g["det_center"])
    ds0, M0, P0, E0 = park.park_streaming_metric(np.zeros(2048, complex))
    assert ds0.size == 0 and M0.size == 0


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_combined(golden, tag):
    from ofdm_sync_math_b200 import combined_sc_min as c
    g = golden(f"combined_{tag}")
    M, P, R = c.minn_streaming_metric(g["rx"])
    Msc, Psc, Rsc = c.schmidl_cox_streaming_metric(g["rx"])
    close(M, g["M"]); close(Msc, g["M_sc"]); close(Psc, g["P_sc"]); close(Rsc, g["R_sc"])
    gate = c.sc_gate_mask(Msc, float(g["sc_gate_threshold"]))
    assert np.array_equal(segments(gate), g["sc_gate"])
    assert c.find_minn_peak(M, smooth_win=int(g["smooth_win"]), gate_mask=gate) == int(g["peak"])
    assert c.find_minn_peak(np.zeros(0)) == 0
    with pytest.raises(ValueError):
        c.find_minn_peak(np.ones(10))
    with pytest.raises(ValueError):
        c.find_minn_peak(np.ones(10), gate_mask=np.zeros(10, bool))
    with pytest.raises(ValueError):
        c.find_minn_peak(np.ones(10), gate_mask=np.zeros(9, bool))


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_zc(golden, tag):
    from ofdm_sync_math_b200 import zc
    g = golden(f"zc_{tag}")
    assert np.max(np.abs(zc.build_pss_symbol(include_cp=False) - g["ref"])) < 1e-12
    corr, peak, start = zc.zc_detect(g["rx"], g["ref"])
    close(corr, g["corr"], 1e-10); assert peak == int(g["peak"]) and start == int(g["start"])


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_zc_v2(golden, tag):
    from ofdm_sync_math_b200 import zc_v2
    g = golden(f"zc_v2_{tag}")
    res = zc_v2.detect_zc_preamble(g["rx"])
    st = res.state
    close(st.corr_mag, g["corr_mag"], 1e-10); close(st.local_sum, g["local_sum"], 1e-10)
    assert np.array_equal(st.metric_valid, g["valid"]); assert np.array_equal(st.above_threshold, g["above"])
    ev = np.array([[e.peak_index, e.gate_start, e.gate_end, e.detected_start] for e in res.events], dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(ev, g["events"])
    close(np.array([e.peak_value for e in res.events]), g["event_values"], 1e-10)
    assert np.array_equal(segments(res.gate_mask), g["gate"])
    # function-level: FSM on the reference's own corr_mag
    st2 = zc_v2.zc_streaming_detection(g["corr_mag"])
    assert np.array_equal(st2.above_threshold, g["above"])
    r2 = zc_v2.detect_zc_peaks(st2, g["ref"].size)
    assert [e.peak_index for e in r2.events] == g["events"][:, 0].tolist()


@pytest.mark.parametrize("tag", ["cir1", "awgn"])
def test_zc_freq(golden, tag):
    from ofdm_sync_math_b200 import zc_freq
    g = golden(f"zc_freq_{tag}")
    bi, tb, te = zc_freq.make_pss_frequency_template()
    assert np.array_equal(bi, g["bin_indices"]) and np.allclose(tb, g["template"], atol=1e-15) and te == float(g["template_energy"])
    m = zc_freq.compute_frequency_metric(g["rx"], bi, tb, te)
    close(m, g["metric"], 1e-10); assert int(np.argmax(m)) == int(g["peak"])
    with pytest.raises(ValueError):
        zc_freq.compute_frequency_metric(g["rx"][:, :2559], bi, tb, te)


@pytest.mark.parametrize("tag", ["cir1", "awgn", "int12"])
def test_minn_rtl(golden, tag):
    from ofdm_sync_math_b200 import minn_rtl as mr
    g = golden(f"minn_rtl_{tag}")
    rx = g["iq"][..., 0].astype(np.float64) + 1j * g["iq"][..., 1].astype(np.float64) if tag == "int12" else g["rx"]
    kw = dict(smooth_shift=int(g["smooth_shift"]), threshold_value=int(g["threshold_value"]),
              threshold_frac_bits=int(g["threshold_frac_bits"]), quarter_len=int(g["quarter_len"]))
    st = mr.minn_rtl_streaming_metric(rx, **kw)
    # the CUDA kernels evaluate the reference's scalar recurrences in its operation order: BIT-EQUAL arrays
    for k in ("corr_total", "corr_positive", "smooth_metric", "energy_total", "corr_scaled", "energy_scaled"):
        assert np.array_equal(getattr(st, k), g[k]), k
    assert np.array_equal(st.metric_valid, g["metric_valid"]); assert np.array_equal(st.above_threshold, g["above"])
    det = mr.detect_minn_rtl(st, hysteresis=int(g["hysteresis"]), timing_offset=int(g["timing_offset"]))
    ev = np.array([[e.peak_index, e.detected_index, *e.gate_segment] for e in det.events], dtype=np.int64).reshape(-1, 4)
    assert np.array_equal(ev, g["events"]); assert np.array_equal(np.asarray(det.gate_segments).reshape(-1, 2), g["gate_segments"])
    assert np.array_equal(segments(det.gate_mask), g["gate_segments"])
    with pytest.raises(ValueError):
        mr.minn_rtl_streaming_metric(rx, smooth_shift=3, threshold_value=1, threshold_frac_bits=15, quarter_len=0)


def test_minn_rtl_integer_datapath(golden):
    """Integer kernel (SV widths, floor-shift smoother) == the CPU integer model bit for bit; == minn_rtl.py for
    corr / energy / valid on the int12 stimulus."""
    from oracle import oracle as orc
    from ofdm_sync_math_b200 import minn_rtl as mr
    g = golden("minn_rtl_int12")
    kw = dict(smooth_shift=3, threshold_value=3276, threshold_frac_bits=15, quarter_len=512)
    for lag_extra in (0, 1):
        st = mr.minn_rtl_int_metric(g["iq"], lag_extra=lag_extra, **kw)
        d = orc.minn_rtl_int(g["iq"], lag_extra=lag_extra, **kw)
        for k in ("corr_total", "corr_positive", "smooth_metric", "energy_total"):
            assert np.array_equal(getattr(st, k), d[k]), (k, lag_extra)
        assert np.array_equal(st.metric_valid, d["metric_valid"]); assert np.array_equal(st.above_threshold, d["above"])
        det = mr.detect_minn_rtl(st, hysteresis=2, timing_offset=0)
        ev_o, seg_o = orc.detect_minn_rtl(d, hysteresis=2, timing_offset=0)
        assert [e.peak_index for e in det.events] == ev_o[:, 0].tolist()
        assert np.array_equal(np.asarray(det.gate_segments).reshape(-1, 2), seg_o)
        if lag_extra == 0:
            assert np.array_equal(st.corr_total, g["corr_total"].astype(np.int64))
            assert np.array_equal(st.energy_total, g["energy_total"].astype(np.int64))


def test_sync_aa_docs_vectors(golden):
    from ofdm_sync_math_b200 import sync_aa
    g = golden("sync_aa_docs")
    for name, csv in (("clean", g["csv_clean"]), ("cfo", g["csv_cfo"])):
        res = sync_aa.aa_detect_streaming(g[f"{name}_rx"], L=512)
        assert np.array_equal(res.state.P, g[f"{name}_P"])                 # reference-order running sums: bit-equal
        close(res.state.R, g[f"{name}_R"], 4e-15); assert np.max(np.abs(res.state.M - g[f"{name}_M"])) <= 4e-15
        assert np.array_equal(res.state.valid, g[f"{name}_valid"])
        assert len(res.events) == 1
        e = res.events[0]
        assert (e.peak_index, e.gate_start, e.gate_end, e.frame_start) == (1523, 1210, 2024, 500)
        assert abs(e.cfo_hz - g[f"{name}_ev_f"][0, 3]) < 1e-9 and abs(e.M_at_peak - 1.0) < 1e-12
        s = csv[:, 0].astype(int)
        assert np.max(np.abs(res.state.M[s] - csv[:, 1])) <= 5.1e-9
        assert np.max(np.abs(res.state.P[s].real - csv[:, 2])) <= 5.1e-3
        # parallel prefix-sum order: same arrays to 1e-12, (the 1523/1524 tie may resolve either way: not asserted)
        r2 = sync_aa.aa_detect_streaming(g[f"{name}_rx"], L=512, order="scan")
        close(r2.state.P, g[f"{name}_P"], 1e-12); close(r2.state.R, g[f"{name}_R"], 1e-12)
        assert r2.events[0].gate_start == 1210 and r2.events[0].gate_end == 2024 and abs(r2.events[0].peak_index - 1523) <= 1


@pytest.mark.parametrize("i", range(5))
def test_sync_aa_grid(golden, i):
    from ofdm_sync_math_b200 import sync_aa
    g = golden(f"sync_aa_grid{i}")
    for order in ("reference", "scan"):
        res = sync_aa.aa_detect_streaming(g["rx"], L=int(g["L"]), order=order)
        if order == "reference":
            assert np.array_equal(res.state.P, g["P"])
        else:
            close(res.state.P, g["P"], 1e-12)
        close(res.state.R, g["R"], 1e-12); assert np.max(np.abs(res.state.M - g["M"])) < 1e-11
        ev_i = np.array([[e.peak_index, e.gate_start, e.gate_end, e.frame_start] for e in res.events], dtype=np.int64).reshape(-1, 4)
        assert np.array_equal(ev_i, g["ev_i"]), order
        ev_f = np.array([[e.P_at_peak.real, e.P_at_peak.imag, e.M_at_peak, e.cfo_hz] for e in res.events]).reshape(-1, 4)
        close(ev_f, g["ev_f"], 1e-10)


def test_detector_cases(golden):
    from ofdm_sync_math_b200 import combined_sc_min as c, minn, sc
    g = golden("detector_cases")
    for k in ("p1", "p2", "p3"):
        assert sc.find_plateau_end_from_metric(g[f"{k}_M"], 512, lookahead=128, smooth_win=16) == int(g[f"{k}_end"]), k
        assert sc.find_plateau_end_from_metric(g[f"{k}_M"], 512) == int(g[f"{k}_end_default"]), k
    kws = dict(g1=dict(smooth_win=16, gate_threshold=0.5), g2=dict(smooth_win=1, gate_threshold=0.3, search_bounds=(800, 1200)),
               g3=dict(smooth_win=8, gate_threshold=0.5, search_bounds=(2000, 100)), g4=dict(smooth_win=4, gate_threshold=0.99))
    for k, kw in kws.items():
        pk, gate, Ms = minn.find_minn_peak(g[f"{k}_M"], **kw)
        assert pk == int(g[f"{k}_peak"]), k
        assert np.array_equal(segments(gate), g[f"{k}_gate"]), k
        close(Ms, g[f"{k}_Ms"], 1e-12)
    assert c.find_minn_peak(g["c1_M"], smooth_win=16, gate_mask=g["c1_gate"]) == int(g["c1_peak"])
    assert c.find_minn_peak(g["c1_M"], smooth_win=16, gate_mask=g["c1_gate"], search_bounds=(850, 2000)) == int(g["c2_peak"])
    assert c._streaming_peak_detector(g["c1_M"], g["c1_gate"]) is not None
