"""GPU parity, SURVEY.md 8(f) rows 1-2: the impairment chain (channel.apply_channel -> core.apply_cfo -> sync_aa.quantize_adc)
and the CP-correlation CFO estimators, against the reference's own outputs (tests/golden/channel_cfo.npz) and the oracle.

Tolerances: float64 path 1e-10 of the signal scale (FFT convolution vs np.convolve), ADC codes equal except where the
float64 sample sits within 1e-9 of a rounding boundary; complex64 path 2e-5 of scale; CFO 1e-6 Hz, indices equal."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu
FS = 30.72e6


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_channel_chain_float64_vs_reference(golden, ci):
    from ofdm_sync_math_b200 import channel, core, engine
    g = golden("channel_cfo")
    tx, cir = g[f"tx{ci}"], g[f"cir{ci}"]
    cir = None if cir.size == 0 else cir
    rx_ref, rc_ref = g[f"rx{ci}"], g[f"rx_cfo{ci}"]
    scale = np.abs(rx_ref).max()
    # drop-in: same seeded generator as the reference run
    rx = channel.apply_channel(tx, float(g[f"snr{ci}"]), np.random.default_rng(int(g[f"noise_seed{ci}"])), channel_impulse_response=cir)
    assert rx.shape == rx_ref.shape and rx.dtype == np.complex128
    assert np.abs(rx - rx_ref).max() <= 1e-10 * scale
    rc = core.apply_cfo(rx_ref, float(g[f"cfo{ci}"]), FS)
    assert np.abs(rc - rc_ref).max() <= 1e-10 * scale
    # fused: FIR + noise + CFO + 12-bit ADC in one call, every branch a stream
    unit = orc.unit_noise_like_reference(int(g[f"noise_seed{ci}"]), rx_ref.shape)
    B = rx_ref.shape[0]
    outs, iqs = [], []
    for b in range(B):
        o, iq = engine.channel_apply(tx.astype(np.complex128), None if cir is None else cir[b], unit_noise=unit[b:b + 1],
                                     snr_db=float(g[f"snr{ci}"]), cfo_hz=float(g[f"cfo{ci}"]), fs=FS,
                                     full_scale=float(g[f"fs_adc{ci}"]), bits=12, want_iq=True)
        outs.append(o.cpu().numpy()[0]); iqs.append(iq.cpu().numpy()[0])
    q_ref, codes_ref = orc.quantize_adc(rc_ref, float(g[f"fs_adc{ci}"]), 12)
    assert np.array_equal(q_ref, g[f"rx_q{ci}"])
    codes = np.stack(iqs)
    diff = codes.astype(np.int64) != codes_ref
    # a code may differ only where the reference sample is within 1e-9 LSB of a .5 boundary
    lsb = float(g[f"fs_adc{ci}"]) / 2048.0
    frac = np.stack((rc_ref.real, rc_ref.imag), axis=-1) / lsb
    near = np.abs(np.abs(frac - np.floor(frac)) - 0.5) < 1e-6
    assert not np.any(diff & ~near), int(diff.sum())
    assert diff.mean() < 1e-4
    assert np.abs(np.stack(outs) - g[f"rx_q{ci}"])[~diff.any(axis=-1)].max() <= 1e-12 * scale


def test_channel_chain_complex64_batch_vs_oracle(golden):
    """Throughput form: complex64, many streams sharing two faded rows, device-resident noise, int16 ingest output."""
    from ofdm_sync_math_b200 import engine
    g = golden("channel_cfo")
    tx = np.stack([g["tx0"], np.roll(g["tx0"], 333)]).astype(np.complex64)
    taps = g["cir0"][1]
    rng = np.random.default_rng(5)
    S = 6
    n_out = tx.shape[1] + taps.size - 1
    unit = (rng.standard_normal((S, n_out)) + 1j * rng.standard_normal((S, n_out))).astype(np.complex64)
    rows = np.array([0, 1, 0, 1, 1, 0], np.int32)
    snr = np.array([0.0, 5.0, 10.0, 15.0, 20.0, 30.0]); cfo = np.array([-9e3, -1e3, 0.0, 500.0, 4e3, 9e3])
    fsc = np.full(S, 3.0)
    out, iq = engine.channel_apply(tx, taps, row_of_stream=rows, unit_noise=unit, snr_db=snr, cfo_hz=cfo, fs=FS, full_scale=fsc,
                                   bits=12, want_iq=True)
    out = out.cpu().numpy(); iq = iq.cpu().numpy()
    for s in range(S):
        ref = orc.channel_apply(tx[rows[s]].astype(np.complex128), taps[None], snr[s], unit[s].astype(np.complex128))[0]
        ref = orc.apply_cfo(ref, cfo[s], FS)
        q, codes = orc.quantize_adc(ref, 3.0, 12)
        assert np.abs(iq[s].astype(np.int64) - codes).max() <= 1            # float32 chain: at most one code off
        assert (iq[s] != codes).mean() < 0.02
        assert np.abs(out[s] - q).max() <= 3.0 / 2048 + 1e-6


@pytest.mark.parametrize("ci", [0, 1, 2])
def test_cp_cfo_estimators_vs_reference(golden, ci):
    from ofdm_sync_math_b200 import core, engine
    g = golden("channel_cfo")
    x = g[f"rx_cfo{ci}"]
    names = ("exact", "early", "edge", "late")
    starts = np.array([int(g[f"est_{k}{ci}"][0]) for k in names])
    for k in names:
        e = g[f"est_{k}{ci}"]
        st = int(e[0])
        assert abs(core.estimate_cfo_from_cp(x, st, 2048, 512, FS) - e[1]) <= 1e-6
        assert abs(core.estimate_cfo_from_cp_robust(x, st, 2048, 512, FS) - e[2]) <= 1e-6
        assert abs(core.estimate_cfo_from_cp_peak(x, st, 2048, 512, FS) - e[3]) <= 1e-6
        c, d = core.estimate_cfo_from_cp_peak_with_index(x, st, 2048, 512, FS, span=100)
        assert abs(c - e[4]) <= 1e-6 and d == int(e[5])
        assert core.find_cp_start_via_corr(x, st, 2048, 512, search_half=300) == int(e[6])
    # batched: the four starts as four frames of the same capture, complex64 input vs the oracle on the cast input
    xb = np.broadcast_to(x.astype(np.complex64), (4,) + x.shape).copy()
    for mode, fn in (("plain", orc.estimate_cfo_from_cp), ("robust", orc.estimate_cfo_from_cp_robust)):
        cfo, bd, P = engine.cp_cfo(torch.as_tensor(xb).cuda(), starts, 2048, 512, FS, mode)
        ref = [fn(xb[f].astype(np.complex128), int(starts[f]), 2048, 512, FS) for f in range(4)]
        assert np.abs(cfo.cpu().numpy() - np.array(ref)).max() <= 1e-6
    cfo, bd, P = engine.cp_cfo(torch.as_tensor(xb).cuda(), starts, 2048, 512, FS, "peak", span=200)
    ref = [orc.estimate_cfo_from_cp_peak_with_index(xb[f].astype(np.complex128), int(starts[f]), 2048, 512, FS, span=200) for f in range(4)]
    assert np.abs(cfo.cpu().numpy() - np.array([r[0] for r in ref])).max() <= 1e-6
    assert bd.cpu().numpy().tolist() == [r[1] for r in ref]
    # windows that leave the capture: NaN / -1 from the batch API, ValueError from the drop-in (the reference raises too)
    cfo, bd, _ = engine.cp_cfo(torch.as_tensor(xb[:1]).cuda(), [x.shape[1] - 1000], 2048, 512, FS, "plain")
    assert np.isnan(cfo.cpu().numpy()[0]) and int(bd.cpu().numpy()[0]) == -1
    with pytest.raises(ValueError):
        core.estimate_cfo_from_cp(x, x.shape[1] - 1000, 2048, 512, FS)


@pytest.mark.parametrize("name", ["sc_cir1", "sc_awgn", "minn_cir1", "minn_awgn"])
def test_rx_chain_vs_reference(golden, name):
    """SURVEY 8f-3 on the GPU (ofs_rx_chain) vs the reference's own run (1 and 2 branches), plus a batch against the oracle."""
    from ofdm_sync_math_b200 import engine
    g = golden(f"rxchain_{name}")
    rx = g["rx"]
    st, cfo = int(g["pilot_cp_start"]), float(g["cfo_est_hz"])
    r = engine.rx_chain(rx[None], st, cfo, g["pilot_used"], g["data_used"])
    h, xh = r["h_est"].cpu().numpy()[0], r["xhat"].cpu().numpy()[0]
    assert np.abs(h - g["h_est"]).max() <= 1e-9 * np.abs(g["h_est"]).max()
    assert np.abs(xh - g["xhat_aligned"]).max() <= 1e-8 * np.abs(g["xhat_aligned"]).max()
    assert abs(complex(r["gain"].cpu().numpy()[0]) - complex(g["gain"])) <= 1e-9 * abs(complex(g["gain"]))
    assert abs(float(r["evm_rms"][0]) - float(g["evm_rms"])) <= 1e-9 and abs(float(r["evm_db"][0]) - float(g["evm_db"])) <= 1e-7
    assert abs(float(r["slope"][0]) - float(g["slope"])) <= 1e-9 and abs(float(r["sto"][0]) - float(g["sto"])) <= 1e-6
    # a batch: shifted starts / other CFO values, complex64 input, per-frame data symbols -> oracle on the same input
    rx32 = rx.astype(np.complex64)
    starts = np.array([st, st - 3, st + 5, st - 40]); cfos = np.array([cfo, 0.0, cfo + 300.0, -cfo])
    du = np.stack([g["data_used"], np.roll(g["data_used"], 1), g["data_used"].conj(), g["data_used"]])
    rb = engine.rx_chain(np.broadcast_to(rx32, (4,) + rx32.shape).copy(), starts, cfos, g["pilot_used"], du)
    for f in range(4):
        o = orc.rx_chain(rx32.astype(np.complex128), int(starts[f]), float(cfos[f]), g["pilot_used"], du[f])
        assert np.abs(rb["h_est"][f].cpu().numpy() - o["h_est"]).max() <= 1e-9 * np.abs(o["h_est"]).max()
        assert abs(float(rb["evm_rms"][f]) - o["evm_rms"]) <= 1e-8 * max(o["evm_rms"], 1.0)
        assert abs(float(rb["sto"][f]) - o["sto"]) <= 1e-6
    # a start outside the capture is flagged (symbols that merely run past the end are zero-padded like np.fft.fft(td, n=N))
    bad = engine.rx_chain(rx32[None], rx.shape[1] + 5, 0.0, g["pilot_used"], g["data_used"])
    assert float(bad["valid"][0]) == 0.0 and np.isnan(float(bad["evm_rms"][0]))
