"""GPU parity over less common parameter combinations of the newer kernels (array, bank, channel, CP-CFO): other FFT sizes,
CP lengths, bin counts, tap counts, branch counts, captures barely longer than one window."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def zadoff_chu(root, length):
    n = np.arange(length)
    return np.exp(-1j * np.pi * root * n * (n + 1) / length)


@pytest.mark.parametrize("n_fft,cp,nbins,n", [(1024, 72, 62, 9000), (512, 0, 24, 3000), (2048, 512, 62, 2700), (2048, 144, 10, 6000),
                                              (256, 32, 62, 5000)])
def test_bank_other_geometries(n_fft, cp, nbins, n):
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(n_fft + nbins)
    half = nbins // 2
    bi = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
    nb = bi.size
    roots = [1, 5, 7]
    T = np.stack([zadoff_chu(r, nb) for r in roots])
    # a capture with template 1 embedded as an OFDM symbol (with CP) at a known offset
    spec = np.zeros(n_fft, complex)
    spec[bi % n_fft] = T[1]
    sym = np.fft.ifft(spec) * np.sqrt(n_fft)
    sym = np.concatenate((sym[n_fft - cp:], sym)) if cp else sym
    x = 0.05 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))
    off = (n - sym.size) // 3
    x[off:off + sym.size] += sym
    x = x.astype(np.complex64)
    bm, bo = engine.zc_bank(x[None], bi, T, n_fft=n_fft, cp=cp)
    bm = bm.cpu().numpy()[0]; bo = bo.cpu().numpy()[0]
    for i in range(len(roots)):
        ref = orc.compute_frequency_metric(x.astype(np.complex128), bi, T[i], float(nb), n_fft=n_fft, cp=cp)
        assert abs(bm[i] - ref.max()) <= 5e-3 * max(ref.max(), 1e-2), (i, bm[i], ref.max())
    ref1 = orc.compute_frequency_metric(x.astype(np.complex128), bi, T[1], float(nb), n_fft=n_fft, cp=cp)
    assert int(np.argmax(bm)) == 1 and int(bo[1]) == int(np.argmax(ref1)) and abs(int(bo[1]) - off) <= 2   # few bins: a broad peak
    m = engine.zc_freq_metric(x[None, None], bi, T[1], float(nb), n_fft=n_fft, cp=cp, out_f64=False, fast=True).cpu().numpy()[0]
    assert m.shape == ref1.shape and np.abs(m - ref1).max() <= 5e-3 * ref1.max()


@pytest.mark.parametrize("L,A,n", [(128, 64, 4096), (256, 64, 8192), (1024, 16, 20480), (512, 33, 1028), (512, 2, 516), (256, 3, 260)])
def test_array_kernel_corner_geometries(L, A, n):
    from ofdm_sync_math_b200 import engine, synth
    cap = synth.aa_capture_host(n, A, seed=L + A, half_len=L, snr_db=8.0, gap=max(64, n // 6))
    r = engine.metric(torch.as_tensor(cap[None]).cuda(), "aa", L, want_pr=True, out_f64=False, path="array")
    ref = orc.aa_detect_streaming(cap.astype(np.complex128), L=L)
    M, P, R = r.M.cpu().numpy()[0], r.P.cpu().numpy()[0], r.R.cpu().numpy()[0]
    err = np.abs(M - ref["M"]) / np.maximum(ref["M"], 1e-6)
    assert err.max() <= 1e-4, (err.max(), int(err.argmax()))
    assert np.abs(P - ref["P"]).max() <= 1e-5 * max(np.abs(ref["P"]).max(), 1e-30)
    assert np.abs(R - ref["R"]).max() <= 1e-5 * ref["R"].max()


@pytest.mark.parametrize("n_taps", [1, 2, 37, 2048])
def test_channel_tap_counts(n_taps):
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(n_taps)
    tx = (rng.standard_normal((2, 5000)) + 1j * rng.standard_normal((2, 5000)))
    taps = (rng.standard_normal(n_taps) + 1j * rng.standard_normal(n_taps)) / np.sqrt(n_taps)
    out, _ = engine.channel_apply(tx, taps)
    out = out.cpu().numpy()
    for r in range(2):
        ref = np.convolve(tx[r], taps, mode="full")
        assert out[r].shape == ref.shape and np.abs(out[r] - ref).max() <= 1e-10 * np.abs(ref).max()


@pytest.mark.parametrize("n_fft,cp,B", [(1024, 72, 1), (256, 64, 3), (2048, 512, 2)])
def test_cp_cfo_other_geometries(n_fft, cp, B):
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(n_fft + B)
    n = 6 * (n_fft + cp)
    x = (rng.standard_normal((4, B, n)) + 1j * rng.standard_normal((4, B, n)))
    # give every frame a real CP structure at a known place and a CFO
    fs = 15.36e6
    for f in range(4):
        st = 700 + 13 * f
        x[f, :, st:st + cp] = x[f, :, st + n_fft:st + n_fft + cp]
        x[f] *= np.exp(1j * 2 * np.pi * (300.0 * (f + 1)) * np.arange(n) / fs)
    starts = np.array([700, 713 + 5, 726 - 9, 739])
    for mode, kw in (("plain", {}), ("robust", {}), ("peak", {"span": 40})):
        cfo, bd, _ = engine.cp_cfo(torch.as_tensor(x).cuda(), starts, n_fft, cp, fs, mode, **kw)
        for f in range(4):
            if mode == "plain":
                ref, rd = orc.estimate_cfo_from_cp(x[f], int(starts[f]), n_fft, cp, fs), int(starts[f])
            elif mode == "robust":
                ref, rd = orc.estimate_cfo_from_cp_robust(x[f], int(starts[f]), n_fft, cp, fs), int(starts[f])
            else:
                ref, rd = orc.estimate_cfo_from_cp_peak_with_index(x[f], int(starts[f]), n_fft, cp, fs, span=40)
            assert abs(float(cfo[f]) - ref) <= 1e-6 * max(1.0, abs(ref)), (mode, f, float(cfo[f]), ref)
            assert int(bd[f]) == rd
    # the peak search lands next to the true CP start of every frame (random data, short CP: not always exactly on it)
    cfo, bd, _ = engine.cp_cfo(torch.as_tensor(x).cuda(), starts, n_fft, cp, fs, "peak", span=40)
    assert np.abs(bd.cpu().numpy() - np.array([700, 713, 726, 739])).max() <= 4
