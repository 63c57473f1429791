"""GPU parity, SURVEY.md 8(f) rank 4 and the wire formats of rank 1.

(1) The 135-case grid of sync_aa.main (run_grid_test, sync_aa.py:829-897) on the batched engine, against the output of the
    unmodified reference run serially (tests/golden/aa_grid.npz, made by oracle/gen_golden.py --only aa_grid).
    Tolerances: detected / num_events / timing_error equal; metric_peak 1e-6 relative and CFO 1e-3 Hz (the FIR runs as an FFT
    overlap-save on the device, 1e-13 of scale away from np.convolve; the detector itself is the bit-equal reference-order
    kernel); clipping percentage and effective bits 1e-9.
(2) HEX24 / AXIS48 12-bit words: bit-exact against docs/preamble_test_vector.hex and the testbench packer's words."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

MAIN_GRID = dict(snr_values=[-5, 0, 5, 10, 15], channels=[None, "cir1", "cir2"], full_scale_ratios=[0.5, 1.0, 2.0],
                 preamble_lengths=[1024, 512, 256], cfo_hz=500.0, plot_samples=False)
CHAN = {"awgn": 0, "cir1": 1, "cir2": 2}


def _check(res, g, rows):
    assert len(res) == len(rows)
    for r, i in zip(res, rows):
        tag = f"case {i}: L={r.preamble_length // 2} {r.channel} snr={r.snr_db} fs={r.full_scale_ratio}"
        assert (r.snr_db, CHAN[r.channel], r.full_scale_ratio, r.preamble_length) == \
            (g["snr_db"][i], g["channel"][i], g["fs_ratio"][i], g["preamble_length"][i]), tag
        assert r.detected == bool(g["detected"][i]), tag
        assert r.num_events == int(g["num_events"][i]), tag
        assert r.timing_error == int(g["timing_error"][i]), tag
        assert abs(r.cfo_estimated_hz - g["cfo_estimated_hz"][i]) <= 1e-3, tag
        assert abs(r.cfo_error_hz - g["cfo_error_hz"][i]) <= 1e-3, tag
        assert abs(r.metric_peak - g["metric_peak"][i]) <= 1e-6 * max(abs(g["metric_peak"][i]), 1e-12), tag
        assert abs(r.clipping_pct - g["clipping_pct"][i]) <= 1e-9, tag
        assert abs(r.effective_bits - g["effective_bits"][i]) <= 1e-9, tag


def test_grid_matches_reference_run(golden):
    from ofdm_sync_math_b200 import sync_aa
    g = golden("aa_grid")
    res = sync_aa.run_grid_test(**MAIN_GRID)
    assert len(res) == 135 and sum(r.detected for r in res) == int(g["detected"].sum())
    _check(res, g, range(135))


@pytest.mark.parametrize("i", [0, 17, 52, 88, 101, 134])
def test_single_case_matches_reference_run(golden, i):
    from ofdm_sync_math_b200 import sync_aa
    g = golden("aa_grid")
    ch = {0: None, 1: "cir1", 2: "cir2"}[int(g["channel"][i])]
    r = sync_aa.run_single_test(float(g["snr_db"][i]), ch, float(g["fs_ratio"][i]), preamble_length=int(g["preamble_length"][i]))
    _check([r], g, [i])


def test_grid_host_helpers_vs_reference(golden):
    from ofdm_sync_math_b200 import sync_aa
    d = golden("sync_aa_docs")
    pre, zc, papr = sync_aa.build_aa_preamble(1024)
    assert np.abs(pre - d["preamble"]).max() <= 1e-14 and zc.size == 300 and 0 < papr < 10
    q = sync_aa.quantize_adc(pre, 2.0)
    assert np.array_equal(q, d["preamble_q12"])
    tone = sync_aa.apply_cfo(d["clean_rx"], 500.0, sync_aa.SAMPLE_RATE_HZ)
    assert np.abs(tone - d["cfo_rx"]).max() <= 1e-13


def test_wire_hex24_vs_reference_vector(golden):
    from ofdm_sync_math_b200 import engine
    d = golden("sync_aa_docs")
    _, codes = orc.quantize_adc(d["preamble"], 2.0)
    words = engine.wire_pack(codes, "hex24")
    assert words.dtype == torch.int32 and np.array_equal(words.cpu().numpy().astype(np.int64), d["hex_preamble"])
    back = engine.wire_unpack(d["hex_preamble"], "hex24")
    assert back.dtype == torch.int16 and np.array_equal(back.cpu().numpy(), codes)


def test_wire_axis48_vs_testbench_packer(golden):
    from ofdm_sync_math_b200 import engine
    g = golden("wire_axis")
    words = engine.wire_pack(g["iq"], "axis48")
    assert words.dtype == torch.int64 and np.array_equal(words.cpu().numpy(), g["words"])
    back = engine.wire_unpack(words, "axis48")
    assert np.array_equal(back.cpu().numpy(), g["iq"])
    assert np.array_equal(orc.wire_pack_axis48(g["iq"]), g["words"])


def test_wire_feeds_the_integer_detector(golden):
    """AXIS words -> int16 IQ -> minn_rtl integer datapath == the same datapath on the original codes."""
    from ofdm_sync_math_b200 import engine
    g = golden("minn_rtl_int12")
    iq = np.ascontiguousarray(g["iq"])
    back = engine.wire_unpack(engine.wire_pack(iq, "axis48"), "axis48")
    assert np.array_equal(back.cpu().numpy(), iq)
    args = (int(g["quarter_len"]), int(g["smooth_shift"]), int(g["threshold_value"]), int(g["threshold_frac_bits"]))
    a = engine.minn_rtl_int(iq, *args)
    b = engine.minn_rtl_int(back, *args)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    assert np.array_equal(b["corr_total"].cpu().numpy()[0], g["corr_total"].astype(np.int64))


def test_wire_rejects_bad_input():
    from ofdm_sync_math_b200 import engine
    with pytest.raises(TypeError):
        engine.wire_pack(np.zeros((4, 2), np.float32), "hex24")
    with pytest.raises(ValueError):
        engine.wire_pack(np.zeros((3, 4, 2), np.int16), "axis48")
    with pytest.raises(ValueError):
        engine.wire_unpack(np.zeros(4, np.int64), "nope")
    assert engine.wire_pack(np.zeros((0, 2), np.int16), "hex24").numel() == 0


# ------------------------------------------------------------------------------------------------- minn / minn_rtl sweeps
SWEEP_KEYS = ("peak", "par", "pmr", "preamble_len", "overhead_pct")


def _check_sweep(got, g, prefix, points):
    """timing_error equal; peak / PAR / PMR 1e-9 relative (float64 metric 1e-12 away from the reference's sequential sums;
    the noise-floor mean is a different summation order)."""
    assert [int(got[p]["timing_error"]) for p in points] == g[f"{prefix}_timing_error"].astype(int).tolist(), prefix
    for k in SWEEP_KEYS:
        ref = g[f"{prefix}_{k}"]
        val = np.array([got[p][k] for p in points], dtype=np.float64)
        assert np.all(np.abs(val - ref) <= 1e-9 * np.maximum(np.abs(ref), 1e-30)), (prefix, k, val, ref)


@pytest.mark.parametrize("ch", [None, "cir1"])
def test_compare_block_lengths_vs_reference_run(golden, ch):
    from ofdm_sync_math_b200 import minn
    g = golden("sweeps")
    Ns = g["block_lengths"].tolist()
    tag = ch or "awgn"
    for snr in (0.0, 10.0):
        r = minn.compare_block_lengths(Ns, ch, snr)
        _check_sweep(r, g, f"block_{tag}_snr{int(snr)}", Ns)
    assert np.abs(r[256]["metric"] - g[f"block_{tag}_metric256"]).max() <= 1e-10 * g[f"block_{tag}_metric256"].max()
    assert r[256]["P_sum"].dtype == np.complex128 and r[256]["R_sum"].shape == r[256]["metric"].shape
    # batched over SNR: the same numbers from one device pass per block length
    both = minn.compare_block_lengths(Ns, ch, [0.0, 10.0])
    for snr in (0.0, 10.0):
        _check_sweep(both[snr], g, f"block_{tag}_snr{int(snr)}", Ns)
    # default SNR is the module's SNR_DB = 0 dB
    _check_sweep(minn.compare_block_lengths(Ns[:1], ch), {k: v[:1] for k, v in g.items() if k.startswith(f"block_{tag}_snr0_")},
                 f"block_{tag}_snr0", Ns[:1])


@pytest.mark.parametrize("ch", [None, "cir1"])
def test_compare_q_values_vs_reference_run(golden, ch):
    from ofdm_sync_math_b200 import minn_rtl
    g = golden("sweeps")
    Qs = g["q_values"].tolist()
    tag = ch or "awgn"
    _check_sweep(minn_rtl.compare_q_values(Qs, ch), g, f"q_{tag}", Qs)
    many = minn_rtl.compare_q_values(Qs, ch, snr_db=[0.0, 20.0])
    _check_sweep(many[0.0], g, f"q_{tag}", Qs)
    assert all(many[20.0][q]["peak"] > 0 for q in Qs)


def test_sweep_tables_print(capsys):
    """The printing front ends of the sweep layer (sync_aa.print_summary_table, minn.run_block_length_comparison,
    minn_rtl.run_q_comparison) run on the device results and show the reference's known answers."""
    from ofdm_sync_math_b200 import minn, minn_rtl, sync_aa
    res = sync_aa.run_grid_test(snr_values=[10], channels=[None, "cir1"], full_scale_ratios=[1.0, 2.0], preamble_lengths=[1024],
                                plot_samples=False)
    sync_aa.print_summary_table(res)
    minn.run_block_length_comparison(None)
    minn_rtl.run_q_comparison(None)
    out = capsys.readouterr().out
    assert "PREAMBLE LENGTH: 1024 samples (L=512)" in out and "AWGN" in out and "CIR1" in out and "(100%)" in out
    assert "BLOCK LENGTH COMPARISON" in out and "Q VALUE COMPARISON" in out
    assert out.count("      +0") >= 4                   # flat AWGN: every block length lands on the expected index (fixture: 0, 0, 0, 0)
