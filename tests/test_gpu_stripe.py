"""GPU parity, part 2: the fast stripe kernel (fp32 products, fp64 carries, TMA ring) and the batched
detectors / sync pipeline against the CPU oracle on seeded synthetic captures.

Tolerances (north_star): metric |dM| <= 1e-4 * max(M, 1e-6); CFO within 1e-5 rad/sample; timing indices equal.
The oracle sees exactly the complex64 / int16 input the GPU sees and computes in float64."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _captures(n_frames, n, kind="sc", seed=0):
    from ofdm_sync_math_b200 import synth
    cirs = synth.load_cirs()
    out = []
    for f in range(n_frames):
        cir = cirs["cir1" if f % 2 == 0 else "cir2"][1] if f % 3 else None
        out.append(synth.tiled_stream_host(n, seed * 1000 + f, kind=kind, cir=cir, snr_db=[0.0, 5.0, 10.0, 20.0][f % 4],
                                           cfo_hz=[-9e3, -1e3, 2e3, 8e3][f % 4]))
    return np.stack(out)


def _dev(x):
    """[frames, L] host array -> [frames, 1 branch, L] device tensor (2-D input would mean (branches, L))."""
    return torch.as_tensor(x).cuda()[:, None]


def _oracle_metric(x, kind, N):
    if kind == "aa":
        return np.stack([orc.aa_detect_streaming(r.astype(np.complex128), L=N)["M"] for r in x])
    k = {"sc": 0, "sc_both": 1, "minn": 2}[kind]
    return np.stack([orc.metric_prefix_c64(r, N, k) for r in x])


def _check_metric(M_gpu, M_ref):
    assert M_gpu.shape == M_ref.shape
    err = np.abs(M_gpu.astype(np.float64) - M_ref) / np.maximum(M_ref, 1e-6)
    worst = float(err.max())
    assert worst <= 1e-4, f"max |dM|/max(M,1e-6) = {worst:.3e} at {np.unravel_index(err.argmax(), err.shape)}"
    return worst


@pytest.mark.parametrize("kind,N", [("sc", 2048), ("sc", 1024), ("sc", 512), ("sc_both", 2048), ("minn", 2048),
                                    ("minn", 4096), ("minn", 1024), ("aa", 512), ("aa", 256), ("aa", 1024)])
@pytest.mark.parametrize("store_mode,tma_mode", [(0, 2), (1, 2), (0, 1)])
def test_stripe_metric_vs_oracle(kind, N, store_mode, tma_mode):
    from ofdm_sync_math_b200 import engine
    n = 40960 + 2 * 4096            # several stripes worth of blocks
    x = _captures(3, n, "minn" if kind == "minn" else "sc", seed=1)
    r = engine.metric(_dev(x), kind, N, want_pr=False, path="stripe", store_mode=store_mode, want_chunk_max=True, tma_mode=tma_mode)
    assert r.path == "stripe"
    M = r.M.cpu().numpy()
    M_ref = _oracle_metric(x, kind, N)
    _check_metric(M, M_ref)
    # chunk maxima: max of M over aligned blocks of 256 causal sample times
    toff = 0 if kind == "aa" else N - 1
    cm = r.chunk_max.cpu().numpy()
    full = np.zeros((x.shape[0], cm.shape[1] * 256), dtype=np.float32)
    full[:, toff:toff + M.shape[1]] = M
    assert np.array_equal(cm, full.reshape(x.shape[0], -1, 256).max(axis=2))


@pytest.mark.parametrize("n", [2048, 2049, 4099, 33333, 65536 + 1023])
def test_stripe_ragged_lengths_and_alignment(n):
    """Odd / short lengths: tail blocks, frames whose rows are not 16-byte aligned (plain-load path), single stripe."""
    from ofdm_sync_math_b200 import engine
    x = _captures(2, n, "sc", seed=2)
    r = engine.metric(_dev(x), "sc", 2048, want_pr=False, path="stripe")
    _check_metric(r.M.cpu().numpy(), _oracle_metric(x, "sc", 2048))


def test_stripe_long_frame_many_stripes():
    from ofdm_sync_math_b200 import engine
    n = 1 << 20
    x = _captures(1, n, "minn", seed=3)
    for kind in ("minn", "sc"):
        r = engine.metric(_dev(x), kind, 2048, want_pr=False, path="stripe")
        _check_metric(r.M.cpu().numpy(), _oracle_metric(x, kind, 2048))


def test_stripe_iq16():
    from ofdm_sync_math_b200 import engine
    x = _captures(2, 50000, "sc", seed=4)
    s = 1500.0 / np.max(np.abs(np.concatenate([x.real, x.imag])))
    iq = np.stack([np.round(x.real * s), np.round(x.imag * s)], axis=-1).astype(np.int16)
    xq = (iq[..., 0] + 1j * iq[..., 1]).astype(np.complex64)
    for kind, N in (("sc", 2048), ("minn", 2048), ("aa", 512)):
        r = engine.metric(torch.as_tensor(iq).cuda()[:, None], kind, N, want_pr=False, path="stripe")
        _check_metric(r.M.cpu().numpy(), _oracle_metric(xq, kind, N))


def test_tile_matches_stripe_and_oracle_c64():
    """The precise tile kernel on complex64 input (float32 outputs) and multi-branch summation."""
    from ofdm_sync_math_b200 import engine
    x = _captures(4, 30000, "sc", seed=5)
    for kind, N in (("sc", 2048), ("sc_both", 1000), ("minn", 2048), ("minn", 900)):
        r = engine.metric(_dev(x), kind, N, want_pr=True, path="tile")
        ref = [getattr(orc, {"sc": "sc_streaming_metric", "sc_both": "schmidl_cox_streaming_metric", "minn": "minn_streaming_metric"}[kind])(
            row.astype(np.complex128), N) for row in x]
        _check_metric(r.M.cpu().numpy(), np.stack([t[0] for t in ref]))
        Pref = np.stack([t[1] for t in ref])
        assert np.max(np.abs(r.P.cpu().numpy() - Pref)) <= 2e-6 * np.max(np.abs(Pref))
    # two branches summed before the metric (sc.py:73-74): frames (2, 2, L)
    xb = x.reshape(2, 2, -1)
    r = engine.metric(torch.as_tensor(xb).cuda(), "sc", 2048, want_pr=True, path="tile")  # (frames, branches, L)
    for f in range(2):
        M0, P0, R0 = orc.sc_streaming_metric(xb[f].astype(np.complex128))
        _check_metric(r.M[f:f + 1].cpu().numpy(), M0[None])


def test_batched_detectors_vs_oracle():
    from ofdm_sync_math_b200 import engine
    x = _captures(6, 30000, "sc", seed=6)
    M = engine.metric(_dev(x), "sc", 2048, want_pr=False, path="stripe").M
    Mh = M.cpu().numpy().astype(np.float64)
    ends = engine.find_plateau_end(M, 512, 128, 16).cpu().numpy()
    assert ends.tolist() == [orc.find_plateau_end_from_metric(r, 512, 128, 16) for r in Mh]
    am = engine.argmax(M).cpu().numpy()
    assert am.tolist() == [int(np.argmax(r)) for r in Mh]
    xm = _captures(5, 30000, "minn", seed=7)
    Mm = engine.metric(_dev(xm), "minn", 2048, want_pr=False, path="stripe").M
    Mmh = Mm.cpu().numpy().astype(np.float64)
    pk, span, _ = engine.find_minn_peak(Mm, 16, 0.5)
    ref = [orc.find_minn_peak(r, 16, 0.5) for r in Mmh]
    assert pk.cpu().numpy().tolist() == [t[0] for t in ref]
    for f, t in enumerate(ref):
        idx = np.flatnonzero(t[1])
        assert span[f].cpu().numpy().tolist() == [idx[0], idx[-1] + 1]
    # S&C-gated Minn peak (combined_sc_min.py:337-365)
    Msc = engine.metric(_dev(xm), "sc_both", 2048, want_pr=False, path="stripe").M
    gate = engine.sc_gate(Msc, 0.6)
    gh = gate.cpu().numpy().astype(bool)
    Msch = Msc.cpu().numpy().astype(np.float64)
    for f in range(xm.shape[0]):
        assert np.array_equal(gh[f], orc.sc_gate(Msch[f], 0.6))
    pg = engine.find_minn_peak_gated(Mm, 16, gate).cpu().numpy()
    assert pg.tolist() == [orc.find_minn_peak_gated(Mmh[f], 16, gh[f]) for f in range(xm.shape[0])]
    # the chunk-max pruned gate must be the same mask, for several thresholds and with an all-zero row (seed rule)
    xz = xm.copy(); xz[2] = 0
    rs = engine.metric(_dev(xz), "sc_both", 2048, want_pr=False, path="stripe", want_chunk_max=True)
    for thr in (0.6, 0.05, 1.0):
        g0 = engine.sc_gate(rs.M, thr).cpu().numpy()
        g1 = engine.sc_gate(rs.M, thr, chunk_max=rs.chunk_max, toff=2047).cpu().numpy()
        assert np.array_equal(g0, g1)
        Mz = rs.M.cpu().numpy().astype(np.float64)
        for f in range(xz.shape[0]):
            assert np.array_equal(g1[f].astype(bool), orc.sc_gate(Mz[f], thr))
        # fused detector (no gate array): same peak as gate + gated peak and as the oracle, same first gate segment
        rm = engine.metric(_dev(xz), "minn", 2048, want_pr=False, path="stripe")
        Mmz = rm.M.cpu().numpy().astype(np.float64)
        for bounds in (None, (5000, 20000), (27000, 27940), (0, 1), (40000, 3)):
            pk2 = engine.find_minn_peak_gated(rm.M, 16, torch.as_tensor(g1, device=rm.M.device), bounds).cpu().numpy()
            pkf, spanf = engine.combined_peak(rm.M, rs.M, rs.chunk_max, 2047, thr, 16, bounds)
            assert pkf.cpu().numpy().tolist() == pk2.tolist(), (thr, bounds)
            for f in range(xz.shape[0]):
                if pk2[f] >= 0:
                    assert pk2[f] == orc.find_minn_peak_gated(Mmz[f], 16, g1[f].astype(bool), bounds), (thr, bounds, f)
                    a, b = spanf[f].cpu().numpy().tolist()
                    assert g1[f][a:b].all() and (b == len(g1[f]) or not g1[f][b] or (bounds and b == min(bounds[1], len(g1[f])))), (thr, bounds, f)


def test_pruned_detectors_equal_unpruned():
    """chunk-max pruning must not change any answer (plateau end, Minn peak / gate span)."""
    from ofdm_sync_math_b200 import engine
    x = _captures(6, 70000, "sc", seed=11)
    r = engine.metric(_dev(x), "sc", 2048, want_pr=False, path="stripe", want_chunk_max=True)
    a = engine.find_plateau_end(r.M, 512, 128, 16)
    b = engine.find_plateau_end(r.M, 512, 128, 16, chunk_max=r.chunk_max, toff=2047)
    assert a.tolist() == b.tolist()
    Mh = r.M.cpu().numpy().astype(np.float64)
    assert a.tolist() == [orc.find_plateau_end_from_metric(m, 512, 128, 16) for m in Mh]
    xm = _captures(6, 70000, "minn", seed=12)
    rm = engine.metric(_dev(xm), "minn", 2048, want_pr=False, path="stripe", want_chunk_max=True)
    for thr in (0.5, 0.05, 0.9):
        p0, s0, _ = engine.find_minn_peak(rm.M, 16, thr)
        p1, s1, _ = engine.find_minn_peak(rm.M, 16, thr, chunk_max=rm.chunk_max, toff=2047)
        assert p0.tolist() == p1.tolist() and s0.tolist() == s1.tolist()
    Mmh = rm.M.cpu().numpy().astype(np.float64)
    p1, s1, _ = engine.find_minn_peak(rm.M, 16, 0.5, chunk_max=rm.chunk_max, toff=2047)
    assert p1.tolist() == [orc.find_minn_peak(m, 16, 0.5)[0] for m in Mmh]


def test_fsm_random_flags_vs_oracle():
    """Gate / hysteresis FSMs on adversarial random flag patterns (dense runs, gaps around the hysteresis)."""
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(8)
    n = 20000
    for trial, hyst in enumerate([0, 1, 2, 5, 64, 256]):
        mag = rng.random((4, n))
        # inject plateaus and ties
        for r in range(4):
            for _ in range(6):
                a = int(rng.integers(3000, n - 600)); mag[r, a:a + int(rng.integers(1, 400))] += 1.0
            mag[r, 5000:5010] = 3.0          # exact ties: first-max vs last-max rules
        W = 2048
        ls, valid, above = engine.zc_streaming_detection(torch.as_tensor(mag).cuda(), W, 40, 15, 0.9)
        evs, gm = engine.zc_events(torch.as_tensor(mag).cuda(), valid, above, 2048, hyst)
        for r in range(4):
            st = orc.zc_streaming_detection(mag[r], W, 40, 15, 0.9)
            assert np.array_equal(above[r].cpu().numpy().astype(bool), st.above_threshold)
            assert np.allclose(ls[r].cpu().numpy(), st.local_sum, rtol=1e-12, atol=1e-9)
            ev_o, val_o, gm_o = orc.detect_zc_peaks(st, 2048, hyst, max_ev=mag.shape[1] + 2)
            e = evs[r]
            got = np.stack([e["peak_index"], e["gate_start"], e["gate_end"], e["aux"]], axis=1) if len(e) else np.zeros((0, 4), np.int64)
            assert np.array_equal(got, ev_o), (trial, r)
            assert np.array_equal(gm[r].cpu().numpy().astype(bool), gm_o)
            # minn_rtl FSM (last max, `>=`) on the same flags
            d = dict(corr_positive=mag[r].copy(), above=st.above_threshold, metric_valid=st.metric_valid)
            ev_r, seg_r = orc.detect_minn_rtl(d, hysteresis=hyst, timing_offset=-7, max_ev=mag.shape[1] + 2)
            er = engine.minn_rtl_events(torch.as_tensor(mag[r]).cuda(), torch.as_tensor(st.metric_valid), torch.as_tensor(st.above_threshold), hyst, -7)[0]
            closed = er[er["closed"] == 1]
            assert np.array_equal(np.stack([closed["peak_index"], closed["aux"], closed["gate_start"], closed["gate_end"]], axis=1)
                                  if len(closed) else np.zeros((0, 4), np.int64), ev_r[:len(closed)]), (trial, r)
            assert np.array_equal(np.stack([er["gate_start"], er["gate_end"]], axis=1) if len(er) else np.zeros((0, 2), np.int64), seg_r)


def test_sync_pipeline_device_and_host(golden):
    """ofs_sync (device) and ofs_sync_host (host buffers, pipelined copies): M, timing, CFO vs oracle."""
    from ofdm_sync_math_b200 import engine
    n = 49152
    x = _captures(5, n, "sc", seed=9)
    Mref = _oracle_metric(x, "sc", 2048)
    ends = [orc.find_plateau_end_from_metric(r, 512, 128, 16) for r in Mref]          # float64 metric of the oracle
    plan = engine.SyncPlan(5, n, "sc", 2048, "c64", cp_len=512, smooth_win=16, sc_delta=16)
    out = plan.run(torch.as_tensor(x).cuda())
    torch.cuda.synchronize()
    rec = out.records_numpy()
    Mg = out.M.cpu().numpy()
    _check_metric(Mg, Mref)
    plan.resolve(torch.as_tensor(x).cuda())
    rec = plan.records_numpy()
    assert rec["timing"].tolist() == ends                 # exact mode (default): equal to the float64 reference, no near-tie allowance
    plain = engine.SyncPlan(5, n, "sc", 2048, "c64", cp_len=512, smooth_win=16, sc_delta=16, exact=False)
    rec32 = plain.run(torch.as_tensor(x).cuda()).records_numpy()
    ends_gpuM = [orc.find_plateau_end_from_metric(r.astype(np.float64), 512, 128, 16) for r in Mg]
    assert rec32["timing"].tolist() == ends_gpuM          # float32-only mode: the detector on the GPU's own metric, exactly
    for f in range(5):
        c = int(rec["coarse"][f]); assert c == max(int(rec["timing"][f]) - 16, 0)
        M0, P0, R0 = orc.metric_prefix_c64(x[f], 2048, 0, want_pr=True)
        cfo_ref = -np.angle(P0[c]) / (2 * np.pi * 1024)
        assert abs(rec["cfo"][f] - cfo_ref) * 2 * np.pi <= 1e-5           # rad / sample
        assert abs(rec["p_re"][f] - P0[c].real) <= 2e-6 * abs(P0[c]) + 1e-6
    # host-buffer path
    xh = torch.as_tensor(x).pin_memory()
    Mh = torch.zeros((5, n - 2047), dtype=torch.float32).pin_memory()
    rh = torch.zeros((5, engine.REC_BYTES), dtype=torch.uint8).pin_memory()
    hs = engine.HostSync()
    rec2 = hs.run(xh, Mh, rh, kind="sc", symbol_len=2048, cp_len=512, smooth_win=16, sc_delta=16)
    assert np.array_equal(Mh.numpy(), Mg)
    assert rec2["timing"].tolist() == rec["timing"].tolist() and np.allclose(rec2["cfo"], rec["cfo"])
    hs.close()
    # the reference-made scale fixture (real channel.py, tests/golden/scale_sc.npz)
    g = golden("scale_sc")
    plan2 = engine.SyncPlan(g["x"].shape[0], g["x"].shape[1], "sc", 2048, "c64")
    o2 = plan2.run(torch.as_tensor(g["x"]).cuda())
    torch.cuda.synchronize()
    _check_metric(o2.M.cpu().numpy(), g["M"])
    assert o2.records_numpy()["timing"].tolist() == g["plateau_end"].tolist()


def test_minn_sync_pipeline():
    from ofdm_sync_math_b200 import engine
    n = 40000
    x = _captures(4, n, "minn", seed=10)
    plan = engine.SyncPlan(4, n, "minn", 2048, "c64", smooth_win=16, gate_threshold=0.5)
    out = plan.run(torch.as_tensor(x).cuda())
    torch.cuda.synchronize()
    rec = out.records_numpy()
    Mg = out.M.cpu().numpy().astype(np.float64)
    assert rec["timing"].tolist() == [orc.find_minn_peak(orc.metric_prefix_c64(r, 2048, 2), 16, 0.5)[0] for r in x]   # float64 oracle
    plain = engine.SyncPlan(4, n, "minn", 2048, "c64", smooth_win=16, gate_threshold=0.5, exact=False)
    rec32 = plain.run(torch.as_tensor(x).cuda()).records_numpy()
    assert rec32["timing"].tolist() == [orc.find_minn_peak(r, 16, 0.5)[0] for r in Mg]
    for f in range(4):
        c = int(rec["coarse"][f])
        M0, P0, R0 = orc.metric_prefix_c64(x[f], 2048, 2, want_pr=True)
        assert abs(rec["cfo"][f] + np.angle(P0[c]) / (2 * np.pi * 512)) * 2 * np.pi <= 1e-5


@pytest.mark.parametrize("kind,N", [("sc", 2048), ("sc_both", 2048), ("minn", 2048), ("aa", 512), ("sc", 512)])
@pytest.mark.parametrize("dtype", ["c64", "iq16"])
def test_stripe_pr_outputs_vs_oracle(kind, N, dtype):
    """The stripe kernel's P / R outputs (path="stripe", want_pr=True): reference-shaped (M, P, R) at HBM speed.
    P, R within 2e-5 of their scale (float32 windows with float64 carries), M to the usual 1e-4."""
    from ofdm_sync_math_b200 import engine
    n = 20480 + 3 * 1024 + 4
    x = _captures(3, n, "minn" if kind == "minn" else "sc", seed=21)
    if dtype == "iq16":
        q = np.round(np.stack((x.real, x.imag), axis=-1) * 300.0).clip(-2047, 2047).astype(np.int16)
        xd = torch.as_tensor(q).cuda()[:, None]
        x = (q[..., 0].astype(np.float32) + 1j * q[..., 1].astype(np.float32)).astype(np.complex64)
    else:
        xd = _dev(x)
    r = engine.metric(xd, kind, N, want_pr=True, path="stripe")
    assert r.path == "stripe" and r.P is not None and r.R is not None
    M, P, R = r.M.cpu().numpy(), r.P.cpu().numpy(), r.R.cpu().numpy()
    for f in range(x.shape[0]):
        if kind == "aa":
            o = orc.aa_detect_streaming(x[f].astype(np.complex128), L=N)
            Mo, Po, Ro = o["M"], o["P"], o["R"]
        else:
            k = {"sc": 0, "sc_both": 1, "minn": 2}[kind]
            Mo, Po, Ro = orc.metric_prefix_c64(x[f], N, k, want_pr=True)
        _check_metric(M[f:f + 1], Mo[None])
        assert np.abs(P[f] - Po).max() <= 2e-5 * np.abs(Po).max()
        assert np.abs(R[f] - Ro).max() <= 2e-5 * Ro.max()


def _oracle_metric_branches(xb, kind, N):
    """Float64 metric of one frame with branches (B, L): P and R summed over the branches before the non-linear step
    (sc.py:73-74, minn.py:106-107, combined_sc_min.py:159-160)."""
    k = {"sc": 0, "sc_both": 1, "minn": 2}[kind]
    P = 0; R = 0
    for b in range(xb.shape[0]):
        _, Pb, Rb = orc.metric_prefix_c64(xb[b], N, k, want_pr=True)
        P = P + Pb; R = R + Rb
    num = np.maximum(P.real, 0.0) ** 2 if kind == "minn" else np.abs(P) ** 2
    return num / np.maximum(R, 1e-12) ** 2


@pytest.mark.parametrize("kind,N", [("sc", 2048), ("sc_both", 2048), ("minn", 2048), ("minn", 4096), ("sc", 512)])
@pytest.mark.parametrize("B", [2, 3, 4])
@pytest.mark.parametrize("dtype", ["c64", "iq16"])
def test_stripe_multibranch_vs_oracle(kind, N, B, dtype):
    """Several branches per frame on the stripe path (MB kernel variant): branch sum before the scan, 1e-4 vs the float64 oracle;
    path="auto" must pick it when the pitches allow the tiled TMA ring."""
    from ofdm_sync_math_b200 import engine
    n = 36864 + 3 * 1024                # multiple of 32: 128-byte branch pitch for both dtypes; several stripes
    F = 3
    x = _captures(F * B, n, "minn" if kind == "minn" else "sc", seed=31).reshape(F, B, n)
    x[1, 1] *= 0.25                     # unequal branch powers
    if dtype == "iq16":
        q = np.round(np.stack((x.real, x.imag), axis=-1) * 300.0).clip(-2047, 2047).astype(np.int16)
        xd = torch.as_tensor(q).cuda()
        x = (q[..., 0].astype(np.float32) + 1j * q[..., 1].astype(np.float32)).astype(np.complex64)
    else:
        xd = torch.as_tensor(x).cuda()
    r = engine.metric(xd, kind, N, want_pr=False, path="auto", want_chunk_max=True)
    assert r.path == "stripe"
    M = r.M.cpu().numpy()
    for f in range(F):
        _check_metric(M[f:f + 1], _oracle_metric_branches(x[f], kind, N)[None])
    toff = N - 1
    cm = r.chunk_max.cpu().numpy()
    full = np.zeros((F, cm.shape[1] * 256), dtype=np.float32)
    full[:, toff:toff + M.shape[1]] = M
    assert np.array_equal(cm, full.reshape(F, -1, 256).max(axis=2))
    # a pitch that is not a multiple of 128 bytes falls back to the precise kernel, same tolerance
    r2 = engine.metric(xd[:, :, :n - 8].contiguous(), kind, N, want_pr=False, path="auto")
    assert r2.path == "tile"


def test_stripe_multibranch_golden_two_branch_fixtures(golden):
    """The reference's own two-branch runs (minn.py:347-351 feeds both CIR channels; combined_sc_min.py likewise) through the
    multi-branch stripe path: complex64 copies of the fixtures' rx, zero-padded to a 128-byte pitch, metric within 1e-4 of the
    reference's float64 arrays."""
    from ofdm_sync_math_b200 import engine
    for name, kinds in (("minn_cir1", (("minn", "M"),)), ("combined_cir1", (("minn", "M"), ("sc_both", "M_sc")))):
        g = golden(name)
        rx = np.asarray(g["rx"])
        assert rx.ndim == 2 and rx.shape[0] == 2
        L = rx.shape[1]
        Lp = (L + 15) // 16 * 16
        x = np.zeros((1, 2, Lp), dtype=np.complex64)
        x[0, :, :L] = rx.astype(np.complex64)
        for kind, key in kinds:
            if key not in g:
                continue
            r = engine.metric(torch.as_tensor(x).cuda(), kind, 2048, want_pr=False, path="auto")
            assert r.path == "stripe"
            Mref = np.asarray(g[key], dtype=np.float64)
            M = r.M.cpu().numpy()[0, :Mref.size].astype(np.float64)
            assert np.max(np.abs(M - Mref) / np.maximum(Mref, 1e-6)) <= 1e-4


def test_sync_pipeline_two_branches_exact():
    """ofs_sync on two-branch frames (stripe MB kernel + exact detector + branch-summed P / CFO record) vs the float64 oracle."""
    from ofdm_sync_math_b200 import engine
    F, B, n = 6, 2, 40960
    x = _captures(F * B, n, "sc", seed=41).reshape(F, B, n)
    plan = engine.SyncPlan(F, n, "sc", 2048, "c64", cp_len=512, smooth_win=16, sc_delta=16, n_branches=B)
    xd = torch.as_tensor(x).cuda()
    plan.run(xd)
    plan.resolve(xd)
    rec = plan.records_numpy()
    for f in range(F):
        Mo = _oracle_metric_branches(x[f], "sc", 2048)
        assert int(rec["timing"][f]) == orc.find_plateau_end_from_metric(Mo, 512, 128, 16)
        c = int(rec["coarse"][f])
        P = sum(orc.metric_prefix_c64(x[f, b], 2048, 0, want_pr=True)[1][c] for b in range(B))
        assert abs(rec["cfo"][f] + np.angle(P) / (2 * np.pi * 1024)) * 2 * np.pi <= 1e-5
