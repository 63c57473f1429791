"""GPU parity: the integer RTL-width Minn datapath (ref/minn_antenna_path.sv:63-194, ref/minn_preamble_detector.sv:247-325) with
the PARALLEL floor-shift smoother against the oracle's serial int64 model -- bit for bit, including the inputs on which
the parallel chains cannot converge and hand over to their neighbour or to the serial kernel."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

KEYS = ("corr_total", "corr_positive", "smooth_metric", "energy_total")


def _check(iq, Q, shift, thr=3276, frac=15, lag_extra=0):
    from ofdm_sync_math_b200 import engine
    d = engine.minn_rtl_int(torch.as_tensor(iq).cuda(), Q, shift, thr, frac, lag_extra)
    for f in range(iq.shape[0]):
        o = orc.minn_rtl_int(iq[f], smooth_shift=shift, threshold_value=thr, threshold_frac_bits=frac, quarter_len=Q, lag_extra=lag_extra)
        for k in KEYS:
            g = d[k][f].cpu().numpy()
            assert np.array_equal(g, o[k]), (k, f, shift, int(np.flatnonzero(g != o[k])[0]))
        assert np.array_equal(d["metric_valid"][f].cpu().numpy().astype(bool), o["metric_valid"])
        assert np.array_equal(d["above_threshold"][f].cpu().numpy().astype(bool), o["above"])


def _frames(F, A, n, seed, amp=2047):
    rng = np.random.default_rng(seed)
    return rng.integers(-amp, amp + 1, size=(F, A, n, 2), dtype=np.int64).astype(np.int16)


@pytest.mark.parametrize("shift", [0, 1, 2, 3, 4, 5])
@pytest.mark.parametrize("n", [1500, 8192, 8193, 40000])
def test_parallel_smoother_bit_exact_random(shift, n):
    _check(_frames(3, 2, n, 10 * shift + n % 7), 256, shift)


def test_parallel_smoother_structured_inputs():
    """Preamble-like bursts between silences and full-scale int16 values: large dynamic range of corr_positive, long runs of
    zeros (the state decays to exactly 0 from above) and a ramp."""
    rng = np.random.default_rng(5)
    F, A, n, Q = 4, 2, 30000, 512
    iq = np.zeros((F, A, n, 2), dtype=np.int16)
    a = rng.integers(-2000, 2001, size=(Q, 2))
    for f in range(F):
        for s in range(3000 + 500 * f, n - 4 * Q, 9000):
            blk = np.concatenate([a, a, -a, -a])
            iq[f, :, s:s + 4 * Q] = blk[None]
        iq[f] += rng.integers(-30, 31, size=(A, n, 2)).astype(np.int16)
    iq[3] = rng.integers(-32768, 32768, size=(A, n, 2), dtype=np.int64).astype(np.int16)     # full int16 range
    for shift in (2, 3, 4):
        _check(iq, Q, shift)
    _check(iq, Q, 3, lag_extra=1)


def test_parallel_smoother_non_converging_input_falls_back():
    """A constant positive corr_positive keeps the two bounding trajectories apart for ever (every state in (c - 2^k, c] is a
    fixed point): the chains cannot converge, the segments are handed over / the frame goes to the serial kernel -- and the
    result is still bit-exact."""
    F, A, n, Q = 2, 1, 20000, 64
    iq = np.zeros((F, A, n, 2), dtype=np.int16)
    iq[..., 0] = 1000                        # constant samples: corr and energy constant once the windows are full
    iq[1, :, 5000:5100, 0] = 1003            # a small disturbance in the second frame
    for shift in (3, 4):
        _check(iq, Q, shift)


def test_parallel_equals_forced_serial_on_bench_shape():
    from ofdm_sync_math_b200 import engine
    iq = torch.as_tensor(_frames(8, 2, 32768, 99)).cuda()
    a = engine.minn_rtl_int(iq, 512, 3, 3276, 15)
    os.environ["OFS_RTL_SERIAL_SMOOTHER"] = "1"
    try:
        b = engine.minn_rtl_int(iq, 512, 3, 3276, 15)
    finally:
        del os.environ["OFS_RTL_SERIAL_SMOOTHER"]
    for k in a:
        assert torch.equal(a[k], b[k]), k
