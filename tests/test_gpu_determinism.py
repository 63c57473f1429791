"""Race hygiene without compute-sanitizer (closed on this GPU pool -- profiles/r2_sanitizer_closed.log): the kernels that
hand data between asynchronous agents (TMA rings + mbarriers, triple-buffered chunk totals, the bank's operand ring and TMEM
accumulators, in-place shared-memory state of the parallel smoother) are run several times on the same input, interleaved
with other work that disturbs scheduling, and must return bit-identical outputs every time.  A missing barrier or a slot
reused too early shows up as a run-to-run difference long before it shows up as a wrong answer."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cplx(rng, *shape):
    return torch.as_tensor((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)).cuda()


def _disturb():
    a = torch.randn(2048, 2048, device="cuda")
    return (a @ a).sum()


def _same(fn, reps=4):
    ref = [t.clone() for t in fn()]
    for _ in range(reps):
        _disturb()
        got = fn()
        for a, b in zip(ref, got):
            assert torch.equal(a, b)


def test_stripe_and_multibranch_ring_deterministic():
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(1)
    x1, x2 = _cplx(rng, 64, 1, 65536), _cplx(rng, 32, 2, 65536)
    for kind in ("sc", "minn", "sc_both"):
        _same(lambda: (lambda r: (r.M, r.chunk_max))(engine.metric(x1, kind, 2048, want_pr=False, path="stripe", want_chunk_max=True)))
        _same(lambda: (lambda r: (r.M, r.chunk_max))(engine.metric(x2, kind, 2048, want_pr=False, path="stripe", want_chunk_max=True)))
    _same(lambda: (lambda r: (r.M, r.P, r.R))(engine.metric(x1, "sc", 2048, want_pr=True, path="stripe")))


def test_sync_exact_mode_and_array_kernel_deterministic():
    from ofdm_sync_math_b200 import engine, synth
    x = synth.make_batch_device(128, 65536, "sc", seed=3)
    plan = engine.SyncPlan(128, 65536, "sc", 2048, "c64")
    _same(lambda: (plan.run(x).records.clone(),))
    rng = np.random.default_rng(2)
    xa = _cplx(rng, 4, 16, 32768)
    aplan = engine.AADetectPlan(4, 16, 32768, 512, 0.15, 128, 15.36e6)

    def run_a():
        aplan.run(xa)
        return aplan.M.clone(), torch.view_as_real(aplan.P).clone(), aplan.mask.clone(), aplan.cnt.clone()
    _same(run_a)


def test_bank_filter_threshold_smoother_deterministic():
    from ofdm_sync_math_b200 import engine
    from ofdm_sync_math_b200.zc import build_pss_symbol, generate_zadoff_chu
    rng = np.random.default_rng(4)
    bi = np.concatenate((np.arange(-31, 0), np.arange(1, 32)))
    T = np.stack([generate_zadoff_chu(r, 62) for r in range(1, 65)])
    xb = _cplx(rng, 8, 16384)
    _same(lambda: engine.zc_bank(xb, bi, T))
    _same(lambda: (engine.zc_freq_metric(xb[:, None], bi, T[24], 62.0, fast="f32"),))
    ref = build_pss_symbol(include_cp=False)
    xz = _cplx(rng, 8, 1, 40000)
    zplan = engine.ZCDetectPlan(8, 1, 40000, ref)

    def run_z():
        zplan.run(xz)
        return zplan.mag.clone(), zplan.mask.clone(), zplan.cnt.clone()
    _same(run_z)
    iq = torch.as_tensor(rng.integers(-2047, 2048, size=(16, 2, 32768, 2)).astype(np.int16)).cuda()
    _same(lambda: tuple(engine.minn_rtl_int(iq, 512, 3, 3276, 15).values()))
