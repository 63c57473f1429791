"""GPU parity of the float32 Park paths (park.py:64-114) against the float64 oracle: the block-FFT kernel (park_fft_kernel: band-limited
self-convolution by 256-point block transforms summed per anti-diagonal in the frequency domain, edge triangles direct) and the
direct O(h) kernel it replaces for h = 128 ... 1024 (OFS_PARK_DIRECT=1)."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _capture(n, seed, nb=1):
    rng = np.random.default_rng(seed)
    x = 0.3 * (rng.standard_normal((nb, n)) + 1j * rng.standard_normal((nb, n)))
    # Park preambles around random centres c: x[c-k] = a[k] e^{i phi}, x[c+k] = e^{i phi} / a[k]  (|a| = 1)
    for b in range(nb):
        for c in rng.integers(2500, max(2501, n - 2500), size=max(1, n // 9000)):
            h = 1024
            a = np.exp(2j * np.pi * rng.random(h))
            ph = np.exp(1j * rng.uniform(0, 2 * np.pi))
            x[b, c - h + 1:c + 1] += a[::-1] * ph
            x[b, c:c + h] += (1.0 / a) * ph          # x[c-k] x[c+k] = ph^2 for every k: a sharp Park peak at c
    return x.astype(np.complex64)


def _check(x, symbol_len, M, P, E, tol_m=1e-4):
    ds, Mo, Po, Eo = orc.park_streaming_metric(x.astype(np.complex128), symbol_len)
    assert M.shape == Mo.shape
    if Mo.size == 0:
        return
    assert np.abs(E - Eo).max() <= 2e-6 * Eo.max()
    assert np.abs(P - Po).max() <= 2e-5 * np.abs(Po).max() + 2e-6 * Eo.max()
    assert np.abs(M - Mo).max() <= tol_m * max(Mo.max(), 1e-6), np.abs(M - Mo).max() / Mo.max()
    assert int(np.argmax(M)) == int(np.argmax(Mo))


@pytest.mark.parametrize("direct", [False, True])
@pytest.mark.parametrize("n", [4099, 12288, 40001])
def test_park_c64_vs_oracle(n, direct, monkeypatch):
    from ofdm_sync_math_b200 import engine
    monkeypatch.setenv("OFS_PARK_DIRECT", "1" if direct else "0")
    x = np.stack([_capture(n, 40 + s)[0] for s in range(3)])
    M, P, E = engine.park_metric(torch.as_tensor(x).cuda()[:, None], 2048)
    for f in range(3):
        _check(x[f][None], 2048, M[f].cpu().numpy(), P[f].cpu().numpy(), E[f].cpu().numpy())


@pytest.mark.parametrize("symbol_len", [256, 768, 1024, 1792, 2048])
def test_park_fft_other_half_lengths(symbol_len, monkeypatch):
    """h = 128 (one block: only the diagonal pair and the edge triangles), 384, 512, 896, 1024."""
    from ofdm_sync_math_b200 import engine
    monkeypatch.setenv("OFS_PARK_DIRECT", "0")
    x = _capture(20000, 7 + symbol_len)
    M, P, E = engine.park_metric(torch.as_tensor(x).cuda()[None], symbol_len)
    _check(x, symbol_len, M[0].cpu().numpy(), P[0].cpu().numpy(), E[0].cpu().numpy())


def test_park_fft_two_branches_and_iq16(monkeypatch):
    from ofdm_sync_math_b200 import engine
    monkeypatch.setenv("OFS_PARK_DIRECT", "0")
    x = _capture(30000, 99, nb=2)
    M, P, E = engine.park_metric(torch.as_tensor(x).cuda()[None], 2048)
    _check(x, 2048, M[0].cpu().numpy(), P[0].cpu().numpy(), E[0].cpu().numpy())
    q = np.round(x[:1] * 400).astype(np.complex64)                       # integer-valued samples: int16 IQ carries them exactly
    iq = np.stack([q.real, q.imag], axis=-1).astype(np.int16)
    M, P, E = engine.park_metric(torch.as_tensor(iq).cuda()[None], 2048)
    _check(q, 2048, M[0].cpu().numpy(), P[0].cpu().numpy(), E[0].cpu().numpy())


def test_park_fft_equals_direct_kernel_on_a_batch(monkeypatch):
    """Same batch through both float32 kernels: metric within 2e-5 of its maximum, identical arg-max per frame."""
    from ofdm_sync_math_b200 import engine
    x = torch.as_tensor(np.stack([_capture(66000, 300 + s)[0] for s in range(8)])).cuda()[:, None]
    monkeypatch.setenv("OFS_PARK_DIRECT", "0")
    Mf, Pf, Ef = engine.park_metric(x, 2048)
    monkeypatch.setenv("OFS_PARK_DIRECT", "1")
    Md, Pd, Ed = engine.park_metric(x, 2048)
    assert float((Mf - Md).abs().max()) <= 2e-5 * float(Md.max())
    assert torch.equal(Mf.argmax(dim=1), Md.argmax(dim=1))
    assert float((Ef - Ed).abs().max()) <= 2e-6 * float(Ed.max())


def test_park_fft_degenerate_inputs(monkeypatch):
    """All-zero frames (M = 0, no NaN), frames shorter than one tile, the shortest frame with an output (L = N + 1), and exact
    invariance of M under scaling by a power of two."""
    from ofdm_sync_math_b200 import engine
    monkeypatch.setenv("OFS_PARK_DIRECT", "0")
    z = torch.zeros((2, 1, 5000), dtype=torch.complex64, device="cuda")
    M, P, E = engine.park_metric(z, 2048)
    assert bool(torch.isfinite(M).all()) and float(M.abs().max()) == 0.0 and float(E.abs().max()) == 0.0
    for n in (2049, 2050, 2100, 3583, 3584, 3585):
        x = _capture(max(n, 5200), n)[:, :n]
        M, P, E = engine.park_metric(torch.as_tensor(x).cuda()[None], 2048)
        assert M.shape[-1] == n - 2048
        ds, Mo, Po, Eo = orc.park_streaming_metric(x.astype(np.complex128), 2048)
        assert np.abs(M[0].cpu().numpy() - Mo).max() <= 1e-4 * max(Mo.max(), 1e-6)
        assert np.abs(E[0].cpu().numpy() - Eo).max() <= 2e-6 * Eo.max()
    x = torch.as_tensor(_capture(30000, 5)).cuda()[None]
    M1, _, _ = engine.park_metric(x, 2048)
    M2, _, _ = engine.park_metric(x * 32.0, 2048)
    assert torch.equal(M1, M2)
