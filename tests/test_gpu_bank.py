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
