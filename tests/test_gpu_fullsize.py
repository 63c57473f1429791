"""GPU parity at BASELINE.json's FULL sizes through size-independent properties (the oracle cannot run 2^30 samples):

* scale invariance: every metric is a ratio |P|^2 / R^2, so M(2x) == M(x) BIT FOR BIT in binary floating point (a factor of
  two changes no rounding) -- a wrong halo, a lost carry or a cross-frame leak anywhere in the batch breaks the equality;
* batch independence: a frame's outputs and detection record do not depend on which other frames share its launch
  (whole batch vs. two halves: identical bits), which also pins the persistent-CTA work distribution at full grid size;
* sampled frames of the full batch (8 of cfg 2's 4096, 4 of cfg 3's 1024) against the float64 oracle: metric and timing index;
* the small-size oracle parity (tests/test_gpu_stripe.py, test_gpu_array.py, test_gpu_bank.py) then carries over."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _need(gb):
    free, _ = torch.cuda.mem_get_info()
    if free < gb * 2 ** 30:
        pytest.skip(f"needs {gb} GB of free HBM")


def _bits(t):
    return t.contiguous().view(torch.int32)


@pytest.mark.parametrize("kind,F,n", [("sc", 4096, 262144), ("minn", 1024, 1 << 20)])
def test_cfg2_cfg3_full_size_properties(kind, F, n):
    """BASELINE cfg 2 (4096 x 262144, S&C) and cfg 3 (1024 x 1M, Minn): metric + detector + CFO."""
    from ofdm_sync_math_b200 import engine, synth
    _need(40)
    x = synth.make_batch_device(F, n, "sc" if kind == "sc" else "minn", seed=77, device=torch.device("cuda", 0), chunk=256 if n > 300000 else 1024)
    plan = engine.SyncPlan(F, n, kind, 2048, "c64", smooth_win=16)
    out = plan.run(x)
    rec = out.records_numpy().copy()
    M1 = out.M.clone()
    # (sc.py normalises by the second-half energy only and minn.py by three quarters: M is not bounded by 1)
    assert bool(torch.isfinite(M1).all()) and float(M1.min()) >= 0.0 and float(M1.max()) < 100.0
    # (0) sampled frames of the FULL batch against the float64 oracle: metric within 1e-4, timing index and coarse index equal,
    #     CFO within 1e-5 rad/sample (the oracle takes ~0.1 s per 262 144-sample frame; 8 random frames out of the batch)
    from oracle import oracle as orc
    rng = np.random.default_rng(2025)
    for f in sorted(rng.choice(F, size=8 if n <= 300000 else 4, replace=False).tolist()):
        xf = x[f].cpu().numpy().astype(np.complex128)
        Mg = M1[f].cpu().numpy()
        if kind == "sc":
            Mo, Po, _ = orc.sc_streaming_metric(xf)
            t_ref = orc.find_plateau_end_from_metric(Mo, 512, 128, 16)
        else:
            Mo, Po, _ = orc.minn_streaming_metric(xf)
            t_ref = orc.find_minn_peak(Mo, 16)[0]
        assert np.max(np.abs(Mg - Mo) / np.maximum(Mo, 1e-6)) <= 1e-4, f
        assert int(rec["timing"][f]) == int(t_ref), (f, rec["timing"][f], t_ref)
    # (1) scale invariance, bit for bit, over the whole batch
    x.mul_(2.0)
    out2 = plan.run(x)
    assert torch.equal(_bits(out2.M), _bits(M1))
    rec2 = out2.records_numpy()
    assert np.array_equal(rec2["timing"], rec["timing"]) and np.array_equal(rec2["coarse"], rec["coarse"])
    assert np.allclose(rec2["cfo"], rec["cfo"], rtol=0, atol=1e-9)
    x.mul_(0.5)
    # (2) batch independence: the two halves launched separately give the same bits and the same records
    h = F // 2
    ph = engine.SyncPlan(h, n, kind, 2048, "c64", smooth_win=16)
    for lo in (0, h):
        oh = ph.run(x[lo:lo + h])
        assert torch.equal(_bits(oh.M), _bits(M1[lo:lo + h]))
        rh = oh.records_numpy()
        assert np.array_equal(rh["timing"], rec["timing"][lo:lo + h])
        assert np.array_equal(rh["cfo"], rec["cfo"][lo:lo + h])
    # detections land on a preamble: frames are 9017 samples long with the preamble N-start at 1337 + 512 (+ CIR delay)
    t = rec["timing"]
    assert (t >= 0).all() and (t < n - 2047).all()
    if kind == "sc":
        phase = (t - (1337 + 512)) % 9017
        assert np.mean((phase < 400) | (phase > 9017 - 64)) > 0.95


def test_cfg5_full_size_properties():
    """BASELINE cfg 5: 64 captures x 64 antennas x 262144 (8.6 GB), fused antenna-array detector."""
    from ofdm_sync_math_b200 import engine, synth
    _need(40)
    F, A, n = 64, 64, 262144
    x = synth.make_batch_device(F * A, n, "sc", seed=5, device=torch.device("cuda", 0)).reshape(F, A, n)
    plan = engine.AADetectPlan(F, A, n, 512, 0.15, 128, 15.36e6, want_r=True)
    plan.run(x)
    M1, P1, R1 = plan.M.clone(), plan.P.clone(), plan.R.clone()
    ev1 = plan.events()
    assert bool(torch.isfinite(M1).all()) and float(M1.max()) <= 1.0 and float(M1.min()) >= 0.0
    x.mul_(2.0)
    plan.run(x)
    assert torch.equal(_bits(plan.M), _bits(M1))                       # ratio: bit-exact
    assert torch.equal(_bits(plan.R), _bits(R1 * 4.0))                 # energies scale by exactly 4
    ev2 = plan.events()
    for a, b in zip(ev1, ev2):
        assert np.array_equal(a["peak_index"], b["peak_index"]) and np.array_equal(a["gate_start"], b["gate_start"])
    x.mul_(0.5)
    ph = engine.AADetectPlan(F // 2, A, n, 512, 0.15, 128, 15.36e6, want_r=True)
    for lo in (0, F // 2):
        ph.run(x[lo:lo + F // 2])
        assert torch.equal(_bits(ph.M), _bits(M1[lo:lo + F // 2]))
        assert torch.equal(_bits(torch.view_as_real(ph.P)), _bits(torch.view_as_real(P1[lo:lo + F // 2])))
        for a, b in zip(ph.events(), ev1[lo:lo + F // 2]):
            assert np.array_equal(a["peak_index"], b["peak_index"]) and np.array_equal(a["gate_end"], b["gate_end"])


def test_cfg4_full_size_bank_batch_independence():
    """BASELINE cfg 4: 2048 captures x 65536, 64 Zadoff-Chu roots on the tensor cores."""
    from ofdm_sync_math_b200 import engine, synth
    from ofdm_sync_math_b200.zc import generate_zadoff_chu
    _need(20)
    F, n = 2048, 65536
    x = synth.make_batch_device(F, n, "sc", seed=9, device=torch.device("cuda", 0))
    bi = np.concatenate((np.arange(-31, 0), np.arange(1, 32)))
    T = np.stack([generate_zadoff_chu(r, 62) for r in range(1, 65)])
    bm, bo = engine.zc_bank(x, bi, T)
    assert bool(torch.isfinite(bm).all()) and float(bm.min()) >= 0.0 and float(bm.max()) <= 1.0 + 1e-3
    assert int(bo.min()) >= 0 and int(bo.max()) < n - 2559
    # the work distribution (captures -> persistent CTAs, segments per capture) differs between these launches
    for lo, hi in ((0, 1024), (1024, 2048), (100, 107)):
        bm2, bo2 = engine.zc_bank(x[lo:hi], bi, T)
        if hi - lo >= 148:                   # same segmentation (one work item per capture): identical bits
            assert torch.equal(bm2, bm[lo:hi]) and torch.equal(bo2, bo[lo:hi])
        else:                                # few captures are cut into more segments: other chain starts, TF32-class differences
            assert float((bm2 - bm[lo:hi]).abs().max()) <= 5e-3
            strong = bm[lo:hi] > 0.2
            assert torch.equal(bo2[strong], bo[lo:hi][strong])
