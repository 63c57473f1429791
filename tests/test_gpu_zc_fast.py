"""GPU parity: the float32 Zadoff-Chu fast path -- 8192-point overlap-save matched filter (zc.py:115-126, zc_v2.py:244-271,
486-498), register-prefix threshold kernel (zc_v2.py:288-336) and the three-launch detect_zc_preamble pipeline
(zc_v2.py:456-516) -- against the float64 oracle on the same complex64 / int16 captures."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

_ORACLE_CACHE = {}


def _oracle_zc_freq(key, x, bi, tb):
    """The float64 oracle metric of a capture, computed once per distinct capture (several tests / parametrisations share them;
    the oracle is the slow part of these tests and runs on the GPU box's host cores)."""
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE[key] = orc.compute_frequency_metric(x.astype(np.complex128), bi, tb, 62.0)
    return _ORACLE_CACHE[key]


def _pss_capture(n, seed, snr_db=10.0, n_pss=3):
    """Noise + a few PSS symbols (with CP) at random offsets, through a short random channel, complex64."""
    from ofdm_sync_math_b200.zc import build_pss_symbol
    rng = np.random.default_rng(seed)
    pss = build_pss_symbol(include_cp=True)
    x = np.zeros(n, complex)
    for p in rng.integers(3000, n - 3000, size=n_pss):
        x[p:p + pss.size] += pss * np.exp(1j * rng.uniform(0, 2 * np.pi))
    h = (rng.standard_normal(8) + 1j * rng.standard_normal(8)) * np.exp(-np.arange(8) / 2.0)
    x = np.convolve(x, h / np.linalg.norm(h))[:n]
    std = np.sqrt(np.mean(np.abs(pss) ** 2) / 10 ** (snr_db / 10) / 2)
    x += std * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    return x.astype(np.complex64)


@pytest.mark.parametrize("n", [14000, 20000, 65536, 70001])
@pytest.mark.parametrize("dtype", ["c64", "iq16"])
def test_matched_filter_8k_blocks_vs_oracle(n, dtype):
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import build_pss_symbol
    ref = build_pss_symbol(include_cp=False)
    x = np.stack([_pss_capture(n, 100 + s) for s in range(2)])
    if dtype == "iq16":
        q = np.round(np.stack((x.real, x.imag), axis=-1) * 400.0).clip(-2047, 2047).astype(np.int16)
        xd = torch.as_tensor(q).cuda()[:, None]
        x = (q[..., 0].astype(np.float32) + 1j * q[..., 1].astype(np.float32)).astype(np.complex64)
    else:
        xd = torch.as_tensor(x).cuda()[:, None]
    ref_norm = np.sqrt(np.sum(np.abs(ref) ** 2))
    for mode in (0, 1, 2):
        corr, mag = engine.zc_matched_filter(xd, ref, mode=mode, out_f64=False)
        _, mag_only = engine.zc_matched_filter(xd, ref, mode=mode, out_f64=False, want_corr=False)
        assert torch.equal(mag, mag_only)
        corr, mag = corr.cpu().numpy(), mag.cpu().numpy()
        for f in range(x.shape[0]):
            c, e = orc.matched_filter_correlation(x[f].astype(np.complex128), ref)
            if mode == 0:
                c = c / (ref_norm * np.sqrt(np.maximum(e, 0.0) + 1e-12))
            elif mode == 1:
                c = c / (ref_norm * np.sqrt(np.maximum(e, 1e-12)))
            scale = np.abs(c).max()
            assert np.abs(corr[f] - c).max() <= 1e-4 * scale, (mode, f, np.abs(corr[f] - c).max() / scale)
            assert np.abs(mag[f] - np.abs(c)).max() <= 1e-4 * scale
            assert int(np.argmax(mag[f])) == int(np.argmax(np.abs(c)))


@pytest.mark.parametrize("n", [30000, 65536])
def test_zc_v2_detect_pipeline_vs_oracle(n):
    """ofs_zc_v2_detect: events equal to the oracle's detect_zc_preamble restatement run on the float64 |corr| of the same
    complex64 capture, and to the oracle's FSM run on the GPU's own float32 |corr|."""
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import build_pss_symbol
    ref = build_pss_symbol(include_cp=False)
    F = 3
    x = np.stack([_pss_capture(n, 200 + s, snr_db=[5.0, 10.0, 20.0][s % 3]) for s in range(F)])
    evs, mag = engine.zc_v2_detect(torch.as_tensor(x).cuda()[:, None], ref)
    mag = mag.cpu().numpy()
    for f in range(F):
        got = [(int(e["peak_index"]), int(e["gate_start"]), int(e["gate_end"]), int(e["aux"])) for e in evs[f]]
        st32 = orc.zc_streaming_detection(mag[f].astype(np.float64), 2048, 64, 15, 0.3)
        ev32, _, _ = orc.detect_zc_peaks(st32, ref.size, 256, max_ev=4096)
        assert got == [tuple(int(v) for v in r) for r in ev32.tolist()], f       # FSM + threshold kernel exact on the GPU's own |corr|
        m64 = orc.zc_v2_corr_mag(x[f].astype(np.complex128), ref, True)
        assert np.abs(mag[f] - m64).max() <= 1e-4 * m64.max()
        st64 = orc.zc_streaming_detection(m64, 2048, 64, 15, 0.3)
        ev64, _, _ = orc.detect_zc_peaks(st64, ref.size, 256, max_ev=4096)
        assert len(got) >= 1 and got == [tuple(int(v) for v in r) for r in ev64.tolist()], f


def test_zc_v2_detect_on_reference_fixture_repeated(golden):
    """The reference's own zc_v2 capture (tests/golden/zc_v2_cir1.npz, 2 branches), tiled to cover several 8192-point blocks:
    every repetition must reproduce the fixture's events (shifted), through the float32 pipeline."""
    from ofdm_sync_math_b200 import engine
    g = golden("zc_v2_cir1")
    rx = np.asarray(g["rx"])
    if rx.ndim == 1:
        rx = rx[None]
    reps = 3
    x = np.tile(rx, (1, reps)).astype(np.complex64)
    evs, mag = engine.zc_v2_detect(torch.as_tensor(x).cuda()[None], g["ref"])
    m64 = orc.zc_v2_corr_mag(x.astype(np.complex128), g["ref"], True)
    st64 = orc.zc_streaming_detection(m64, 2048, 64, 15, 0.3)
    ev64, _, _ = orc.detect_zc_peaks(st64, g["ref"].size, 256, max_ev=4096)
    got = [(int(e["peak_index"]), int(e["gate_start"]), int(e["gate_end"]), int(e["aux"])) for e in evs[0]]
    assert got == [tuple(int(v) for v in r) for r in ev64.tolist()]
    # first repetition = the fixture itself (the tail of the capture changes nothing before its own events)
    n_first = int((g["events"][:, 2] < rx.shape[1]).sum())
    assert [t[0] for t in got[:n_first]] == g["events"][:n_first, 0].tolist()


@pytest.mark.parametrize("n", [4000, 20000, 65536])
def test_zc_freq_f32_sliding_dft_vs_oracle(n):
    """compute_frequency_metric (zc_freq.py:62-99) on the packed-fp32 sliding-DFT kernel: float32 metric within 1e-4 of the
    float64 oracle's maximum (north_star float tolerance) on noisy captures with PSS symbols, silence and a loud burst."""
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import generate_zadoff_chu
    half = 31
    bi = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
    tb = generate_zadoff_chu(25, 62)
    x = np.stack([_pss_capture(n, 300 + s, snr_db=[0.0, 10.0, 25.0][s % 3], n_pss=2 if n > 8000 else 1) for s in range(3)]) \
        if n > 8000 else np.stack([_pss_capture(8000, 300 + s)[:n] for s in range(3)])
    x[1, n // 2:n // 2 + 300] *= 40.0            # a loud burst: the recurrence must forget it once it has left the window
    x[2, : n // 4] = 0                           # exact silence
    m = engine.zc_freq_metric(torch.as_tensor(x).cuda()[:, None], bi, tb, 62.0, fast="f32").cpu().numpy()
    for f in range(x.shape[0]):
        mo = _oracle_zc_freq(("burst", n, f), x[f], bi, tb)
        assert m[f].shape == mo.shape
        err = np.abs(m[f] - mo).max()
        assert err <= 1e-4 * mo.max(), (f, err / mo.max())
        assert int(np.argmax(m[f])) == int(np.argmax(mo))


def test_zc_freq_f32_on_reference_fixture(golden):
    from ofdm_sync_math_b200 import engine
    for tag in ("cir1", "awgn"):
        g = golden(f"zc_freq_{tag}")
        rx = np.asarray(g["rx"])
        rx = rx[0] if rx.ndim == 2 else rx          # the fast kernel takes one branch
        x = rx.astype(np.complex64)
        m = engine.zc_freq_metric(torch.as_tensor(x).cuda()[None, None], g["bin_indices"], g["template"], float(g["template_energy"]),
                                  fast="f32").cpu().numpy()[0]
        mo = orc.compute_frequency_metric(x.astype(np.complex128), g["bin_indices"], g["template"], float(g["template_energy"]))
        assert np.abs(m - mo).max() <= 1e-4 * mo.max()
        assert int(np.argmax(m)) == int(np.argmax(mo))


@pytest.mark.parametrize("bpi", [None, "1", "64"])
@pytest.mark.parametrize("n", [4000, 20000, 65536])
def test_zc_freq_fft_form_vs_oracle(n, bpi, monkeypatch):
    """compute_frequency_metric (zc_freq.py:62-99) in FFT form (ofs_zc_freq_metric_fft): float32 metric within 1e-4 of the float64
    oracle's maximum on noisy captures with PSS symbols, exact silence and a loud burst -- with every block anchored on its own
    (blocks per item = 1), with one anchor per capture and the float64 carry of E across all blocks (64), and the default split."""
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import generate_zadoff_chu
    if bpi is None:
        monkeypatch.delenv("OFS_ZQF_BLOCKS_PER_ITEM", raising=False)
    else:
        monkeypatch.setenv("OFS_ZQF_BLOCKS_PER_ITEM", bpi)
    half = 31
    bi = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
    tb = generate_zadoff_chu(25, 62)
    x = np.stack([_pss_capture(n, 300 + s, snr_db=[0.0, 10.0, 25.0][s % 3], n_pss=2 if n > 8000 else 1) for s in range(3)]) \
        if n > 8000 else np.stack([_pss_capture(8000, 300 + s)[:n] for s in range(3)])
    x[1, n // 2:n // 2 + 300] *= 40.0            # a loud burst: E must come back down by cancellation
    x[2, : n // 4] = 0                           # exact silence: E = 0, the FFT's rounding residue must not show
    m = engine.zc_freq_metric(torch.as_tensor(x).cuda()[:, None], bi, tb, 62.0, fast="fft").cpu().numpy()
    for f in range(x.shape[0]):
        mo = _oracle_zc_freq(("burst", n, f), x[f], bi, tb)
        assert m[f].shape == mo.shape
        err = np.abs(m[f] - mo).max()
        assert err <= 1e-4 * mo.max(), (f, err / mo.max(), int(np.argmax(np.abs(m[f] - mo))))
        assert int(np.argmax(m[f])) == int(np.argmax(mo))


def test_zc_freq_fft_form_on_reference_fixture(golden):
    """The reference script's own input, ALL of its receive branches (zc_freq.py:88-97 sums them), as complex64."""
    from ofdm_sync_math_b200 import engine
    for tag in ("cir1", "awgn"):
        g = golden(f"zc_freq_{tag}")
        rx = np.atleast_2d(np.asarray(g["rx"]))
        x = rx.astype(np.complex64)
        m = engine.zc_freq_metric(torch.as_tensor(x).cuda()[None], g["bin_indices"], g["template"], float(g["template_energy"]),
                                  fast="fft").cpu().numpy()[0]
        mo = orc.compute_frequency_metric(x.astype(np.complex128), g["bin_indices"], g["template"], float(g["template_energy"]))
        assert m.shape == mo.shape
        assert np.abs(m - mo).max() <= 1e-4 * mo.max(), (tag, rx.shape, np.abs(m - mo).max() / mo.max())
        assert int(np.argmax(m)) == int(np.argmax(mo))


@pytest.mark.parametrize("nb", [2, 3])
def test_zc_freq_fft_form_branches_summed(nb, monkeypatch):
    """Several receive branches per capture: correlations and in-band energies summed over the branches before the ratio
    (zc_freq.py:88-97) -- float32 metric within 1e-4 of the float64 oracle's maximum, carry across blocks forced."""
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import generate_zadoff_chu
    monkeypatch.setenv("OFS_ZQF_BLOCKS_PER_ITEM", "64")
    n = 30000
    bi = np.concatenate((np.arange(-31, 0), np.arange(1, 32)))
    tb = generate_zadoff_chu(25, 62)
    x = np.stack([np.stack([_pss_capture(n, 500 + 10 * f + b, snr_db=[0.0, 10.0][f % 2], n_pss=2) for b in range(nb)]) for f in range(2)])
    x[1, 0, n // 2:n // 2 + 300] *= 40.0          # a burst on one branch only
    m = engine.zc_freq_metric(torch.as_tensor(x).cuda(), bi, tb, 62.0, fast="fft").cpu().numpy()
    for f in range(2):
        mo = orc.compute_frequency_metric(x[f].astype(np.complex128), bi, tb, 62.0)
        err = np.abs(m[f] - mo).max()
        assert err <= 1e-4 * mo.max(), (f, err / mo.max(), int(np.argmax(np.abs(m[f] - mo))))
        assert int(np.argmax(m[f])) == int(np.argmax(mo))


@pytest.mark.parametrize("n_fft,cp,nb", [(2048, 512, 62), (1024, 72, 40), (512, 0, 12), (1536, 100, 62)])
def test_zc_freq_fft_form_other_geometries(n_fft, cp, nb):
    """Other FFT sizes (power of two: table twiddles in the anchor; 1536: sincospi), prefix lengths and bin counts."""
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(n_fft + nb)
    n = 30000
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    bi = np.concatenate((np.arange(-(nb // 2), 0), np.arange(1, nb // 2 + 1)))
    tb = np.exp(2j * np.pi * rng.random(nb))
    # a preamble made of exactly these bins
    spec = np.zeros(n_fft, complex); spec[bi % n_fft] = tb
    x[7000:7000 + n_fft] += (8.0 * np.fft.ifft(spec) * np.sqrt(n_fft)).astype(np.complex64)
    m = engine.zc_freq_metric(torch.as_tensor(x).cuda()[None, None], bi, tb, float(nb), n_fft=n_fft, cp=cp, fast="fft").cpu().numpy()[0]
    # oracle for general geometry: direct sliding DFT in float64
    n_off = n - (n_fft + cp) + 1
    xd = x.astype(np.complex128)
    k = bi % n_fft
    W = np.exp(-2j * np.pi * np.outer(k, np.arange(n_fft)) / n_fft)
    offs = np.unique(np.concatenate((np.arange(0, n_off, 97), np.arange(7000 - cp - 20, 7000 - cp + 20), [n_off - 1])))
    for o in offs:
        bins = W @ xd[o + cp:o + cp + n_fft]
        ref = abs(np.vdot(tb, bins)) ** 2 / max(float(nb) * np.sum(np.abs(bins) ** 2), 1e-12)
        assert abs(m[o] - ref) <= 1e-4, (o, m[o], ref)
    assert int(np.argmax(m)) == 7000 - cp


@pytest.mark.parametrize("n", [2560, 2561, 2559 + 6145, 2559 + 6146, 2559 + 2 * 6145, 2559 + 2 * 6145 + 1])
def test_zc_freq_fft_form_lengths_around_block_boundaries(n):
    """One offset, one more, exactly one 6145-offset block, one offset into the second block, exactly two blocks, ...: the
    last valid offset of a block reads sample 8192 of the block (the first of the next one), the tail block is ragged."""
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import generate_zadoff_chu
    bi = np.concatenate((np.arange(-31, 0), np.arange(1, 32)))
    tb = generate_zadoff_chu(25, 62)
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) + 1j * rng.standard_normal(n)).astype(np.complex64)
    m = engine.zc_freq_metric(torch.as_tensor(x).cuda()[None, None], bi, tb, 62.0, fast="fft").cpu().numpy()[0]
    mo = orc.compute_frequency_metric(x.astype(np.complex128), bi, tb, 62.0)
    assert m.shape == mo.shape == (n - 2559,)
    assert np.abs(m - mo).max() <= 1e-4 * mo.max()


def test_zc_freq_fft_form_degenerate_inputs():
    """All-zero capture -> metric 0 everywhere (zc_freq.py:96: 0 / eps), no NaN; one bin; scaling the capture by a power of two
    changes no bit of the metric (a ratio of two quadratic forms, float32 arithmetic is exact under such scaling)."""
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import generate_zadoff_chu
    bi = np.concatenate((np.arange(-31, 0), np.arange(1, 32)))
    tb = generate_zadoff_chu(25, 62)
    z = torch.zeros((1, 1, 20000), dtype=torch.complex64, device="cuda")
    m0 = engine.zc_freq_metric(z, bi, tb, 62.0, fast="fft")
    assert bool(torch.isfinite(m0).all()) and float(m0.abs().max()) == 0.0
    x = _pss_capture(20000, 77)
    xt = torch.as_tensor(x).cuda()[None, None]
    m1 = engine.zc_freq_metric(xt, bi, tb, 62.0, fast="fft")
    m2 = engine.zc_freq_metric(xt * 64.0, bi, tb, 62.0, fast="fft")
    assert torch.equal(m1, m2)
    # few bins: E = sum |bins|^2 over 8 bins dips far below its mean now and then, and the recurrence carries ~3e-8 of the recent
    # maximum of E as absolute error -- 1e-4 wherever E is within 20 dB of its median, 1e-2 at the dips (the 62-bin sum of the
    # reference's geometry never dips; the float64-prefix path has no such limit)
    n = 6000
    x8 = x[:n]
    k = np.array([-4, -3, -2, -1, 1, 2, 3, 4])
    t8 = np.exp(2j * np.pi * np.random.default_rng(3).random(8))
    m3 = engine.zc_freq_metric(torch.as_tensor(x8).cuda()[None, None], k, t8, 8.0, fast="fft").cpu().numpy()[0]
    W = np.exp(-2j * np.pi * np.outer(k % 2048, np.arange(2048)) / 2048)
    win = np.lib.stride_tricks.sliding_window_view(x8.astype(np.complex128)[512:], 2048)      # window of offset o starts at o + cp
    bins = win @ W.T                                                                            # [offsets, 8]
    E = np.sum(np.abs(bins) ** 2, axis=1)
    ref = np.abs(bins @ np.conj(t8)) ** 2 / np.maximum(8.0 * E, 1e-12)
    assert m3.shape == ref.shape
    ok = E >= 0.01 * np.median(E)
    assert np.abs(m3 - ref)[ok].max() <= 1e-4 * ref.max()
    assert np.abs(m3 - ref).max() <= 1e-2 * ref.max()


def test_zc_freq_fft_form_int16_iq_two_branches():
    """int16-IQ captures [F, B, n, 2] (the 12-bit ADC format) through the FFT form: same metric as the oracle on the same integers."""
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import generate_zadoff_chu
    bi = np.concatenate((np.arange(-31, 0), np.arange(1, 32)))
    tb = generate_zadoff_chu(25, 62)
    n = 20000
    x = np.stack([_pss_capture(n, 800 + b, snr_db=10.0, n_pss=2) for b in range(2)])
    q = np.round(x / np.abs(x).max() * 2000.0)                      # integer-valued I / Q within 12 bits
    q = (q.real + 1j * q.imag)
    iq = np.stack([q.real, q.imag], axis=-1).astype(np.int16)       # [2, n, 2]
    m = engine.zc_freq_metric(torch.as_tensor(iq).cuda()[None], bi, tb, 62.0, fast="fft").cpu().numpy()[0]
    mo = orc.compute_frequency_metric(q.astype(np.complex128), bi, tb, 62.0)
    assert m.shape == mo.shape
    assert np.abs(m - mo).max() <= 1e-4 * mo.max()
    assert int(np.argmax(m)) == int(np.argmax(mo))
