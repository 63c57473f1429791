"""GPU parity of the tcgen05 correlator bank (K6) against the CPU oracle's zc_freq metric, root by root.
Tolerance (TF32 operands, FP32 accumulation): |d metric| <= 5e-3 * max(metric); the arg-max offset of the
transmitted root must be identical."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def zadoff_chu(root, length=62):
    n = np.arange(length)
    return np.exp(-1j * np.pi * root * n * (n + 1) / length)


@pytest.mark.parametrize("tag", ["awgn", "cir1"])
def test_bank_vs_oracle(golden, tag):
    from ofdm_sync_math_b200 import engine
    g = golden(f"zc_freq_{tag}")
    rx = g["rx"][0].astype(np.complex64)
    rx2 = np.stack([rx, np.roll(rx, 777)])                      # two captures
    roots = np.arange(1, 65)
    T = np.stack([zadoff_chu(r) for r in roots])
    bm, bo = engine.zc_bank(rx2, g["bin_indices"], T)
    bm = bm.cpu().numpy(); bo = bo.cpu().numpy()
    assert bm.shape == (2, 64) and bo.shape == (2, 64)
    for cap, x in enumerate(rx2):
        for r in (25, 1, 64, 31):
            m = orc.compute_frequency_metric(x.astype(np.complex128), g["bin_indices"], zadoff_chu(r), 62.0)
            assert abs(bm[cap, r - 1] - m.max()) <= 5e-3 * max(m.max(), 1e-3), (cap, r, bm[cap, r - 1], m.max())
            if r == 25:
                assert int(bo[cap, r - 1]) == int(np.argmax(m))
        # the transmitted root (25) wins the bank
        assert int(np.argmax(bm[cap])) == 24
    # the second capture is the first one rolled by 777 samples (one branch only here: the 2-branch reference peak differs)
    assert int(bo[1, 24]) == int(bo[0, 24]) + 777
    if tag == "awgn":
        assert int(bo[0, 24]) == int(g["peak"])


@pytest.mark.parametrize("tag", ["awgn", "cir1"])
def test_zc_freq_fast_path_vs_oracle(golden, tag):
    """zc_freq metric row through the bank kernel (ofs_zc_freq_metric_fast): |d metric| <= 5e-3 * max, same arg-max."""
    from ofdm_sync_math_b200 import engine
    g = golden(f"zc_freq_{tag}")
    rx = g["rx"][0].astype(np.complex64)
    caps = np.stack([rx, np.roll(rx, -1234), rx[::-1].copy()])
    tb = zadoff_chu(25)
    m = engine.zc_freq_metric(caps[:, None, :], g["bin_indices"], tb, 62.0, out_f64=False, fast=True).cpu().numpy()
    for c, x in enumerate(caps):
        ref = orc.compute_frequency_metric(x.astype(np.complex128), g["bin_indices"], tb, 62.0)
        assert m[c].shape == ref.shape
        assert np.abs(m[c] - ref).max() <= 5e-3 * ref.max(), (c, np.abs(m[c] - ref).max(), ref.max())
        if c < 2:
            assert int(np.argmax(m[c])) == int(np.argmax(ref))


def test_bank_more_than_64_roots_and_segments(golden):
    """128 templates = two passes of 64; a long capture is cut into several work items per capture."""
    from ofdm_sync_math_b200 import engine
    g = golden("zc_freq_awgn")
    rx = g["rx"][0].astype(np.complex64)
    rng = np.random.default_rng(3)
    noise = lambda k: 0.3 * (rng.standard_normal(k) + 1j * rng.standard_normal(k)).astype(np.complex64)
    long_cap = np.concatenate([noise(20000), rx, noise(12000)])
    T = np.stack([zadoff_chu(i % 61 + 1) * (1.0 + 0.01 * i) for i in range(128)])      # template i = root i % 61 + 1, own scale
    bm, bo = engine.zc_bank(long_cap[None], g["bin_indices"], T)
    bm = bm.cpu().numpy()[0]; bo = bo.cpu().numpy()[0]
    assert bm.shape == (128,)
    for i in (24, 85, 127):                   # 24 and 85 are both root 25, the transmitted one (one per pass)
        ref = orc.compute_frequency_metric(long_cap.astype(np.complex128), g["bin_indices"], T[i], float(np.sum(np.abs(T[i]) ** 2)))
        assert abs(bm[i] - ref.max()) <= 5e-3 * max(ref.max(), 1e-3), (i, bm[i], ref.max())
        if i != 127:
            assert int(bo[i]) == int(np.argmax(ref)) == 20000 + int(g["peak"])
    assert set(np.argsort(bm)[-2:].tolist()) == {24, 85}
