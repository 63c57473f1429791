"""GPU parity, part 4: the antenna-array kernel (metric_array.cu) and the fused detector ofs_aa_detect against the CPU
oracle of sync_aa.aa_detect_streaming (sync_aa.py:421-571).

Tolerances (north_star): metric |dM| <= 1e-4 * max(M, 1e-6); P, R within 1e-5 of their scale; event indices equal; CFO
within 1e-5 rad/sample.  The oracle sees exactly the complex64 / int16 input the GPU sees and computes in float64."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _as_c128(cap):
    return (cap[..., 0].astype(np.float64) + 1j * cap[..., 1].astype(np.float64)) if cap.dtype == np.int16 else cap.astype(np.complex128)


def _check(M, P, R, ref, n):
    Mo, Po, Ro = ref["M"], ref["P"], ref["R"]
    err = np.abs(M.astype(np.float64) - Mo) / np.maximum(Mo, 1e-6)
    assert err.max() <= 1e-4, f"max |dM|/max(M,1e-6) = {err.max():.3e} at {err.argmax()} of {n}"
    assert np.abs(P - Po).max() <= 1e-5 * max(np.abs(Po).max(), 1e-30)
    assert np.abs(R - Ro).max() <= 1e-5 * max(Ro.max(), 1e-30)


@pytest.mark.parametrize("L", [128, 256, 512, 1024])
@pytest.mark.parametrize("A,n,int12", [(1, 2024, False), (2, 40960, False), (5, 33332, False), (64, 16384, False), (4, 20000, True),
                                       (3, 1500, False), (2, 100, False)])
def test_array_metric_vs_oracle(L, A, n, int12):
    from ofdm_sync_math_b200 import engine, synth
    caps = [synth.aa_capture_host(n, A, seed=100 * L + A + f, half_len=L, snr_db=[0.0, 10.0, 25.0][f % 3], int12=int12) for f in range(3)]
    x = torch.as_tensor(np.stack(caps)).cuda()
    r = engine.metric(x, "aa", L, want_pr=True, out_f64=False, path="array")
    assert r.path == "array"
    M, P, R = r.M.cpu().numpy(), r.P.cpu().numpy(), r.R.cpu().numpy()
    assert M.shape == (3, n)
    for f in range(3):
        ref = orc.aa_detect_streaming(_as_c128(caps[f]), L=L)
        _check(M[f], P[f], R[f], ref, n)


def test_array_auto_path_and_tile_agree():
    from ofdm_sync_math_b200 import engine, synth
    caps = np.stack([synth.aa_capture_host(30000, 8, seed=f, half_len=512) for f in range(2)])
    x = torch.as_tensor(caps).cuda()
    a = engine.metric(x, "aa", 512, want_pr=True, out_f64=False)          # auto -> array (branches >= 2)
    t = engine.metric(x, "aa", 512, want_pr=True, out_f64=False, path="tile")
    assert a.path == "array" and t.path == "tile"
    assert torch.allclose(a.M, t.M, rtol=0, atol=2e-6)
    assert torch.allclose(torch.view_as_real(a.P), torch.view_as_real(t.P), rtol=0, atol=1e-5 * float(t.P.abs().max()))
    # odd sample counts are not 16-byte rows: auto falls back to the tile kernel
    o = engine.metric(x[:, :, :29999], "aa", 512, want_pr=True, out_f64=False)
    assert o.path == "tile"


@pytest.mark.parametrize("A,int12", [(1, False), (4, False), (64, False), (8, True)])
def test_fused_aa_detect_vs_oracle(A, int12):
    """ofs_aa_detect: events from the fused bitmask path == oracle FSM (sync_aa.py:495-568) on the GPU's own M / P, and ==
    the all-float64 oracle on these captures (no sample sits within float32 rounding of the 0.15 threshold)."""
    from ofdm_sync_math_b200 import engine, synth
    F, n, L = 3, 36864 + 4 * (A % 3), 512
    caps = [synth.aa_capture_host(n, A, seed=7 + f + A, half_len=L, snr_db=[5.0, 12.0, 30.0][f], cfo_hz=[500.0, -2000.0, 4000.0][f],
                                  int12=int12) for f in range(F)]
    x = torch.as_tensor(np.stack(caps)).cuda()
    plan = engine.AADetectPlan(F, A, n, L, 0.15, 128, 15.36e6, in_dtype="iq16" if int12 else "c64", want_r=True)
    plan.run(x)
    evs = plan.events()
    M, P = plan.M.cpu().numpy(), plan.P.cpu().numpy()
    for f in range(F):
        ref = orc.aa_detect_streaming(_as_c128(caps[f]), L=L, threshold=0.15, hysteresis=128)
        _check(M[f], P[f], plan.R[f].cpu().numpy(), ref, n)
        e = evs[f]
        assert len(e) >= 5
        got = np.stack([e["peak_index"], e["gate_start"], e["gate_end"], e["aux"]], axis=1)
        # (1) oracle FSM on the GPU's arrays: exact
        ev_i = np.zeros((64, 4), np.int64); ev_f = np.zeros((64, 4))
        P64 = np.ascontiguousarray(P[f].astype(np.complex128)); M64 = M[f].astype(np.float64)
        import ctypes as C
        nev = orc.lib().orc_aa_events(orc._p(P64.view(np.float64)), orc._p(M64), orc._p(ref["valid"].astype(np.uint8)), C.c_int64(n),
                                      C.c_int64(L), C.c_double(0.15), C.c_int64(128), C.c_double(15.36e6), orc._p(ev_i), orc._p(ev_f),
                                      C.c_int64(64))
        assert np.array_equal(got, ev_i[:nev])
        # (2) the float64 oracle end to end
        assert np.array_equal(got[:, 1:3], ref["ev_i"][:, 1:3])
        assert np.array_equal(got[:, 0], ref["ev_i"][:, 0])
        cfo_err = np.abs(e["cfo"] - ref["ev_f"][:, 3]) * 2 * np.pi / 15.36e6
        assert cfo_err.max() <= 1e-5


def test_pipelined_record_gather_single_rank_nccl():
    """dist.PipelinedGatherer (records of step k all-gathered on a side stream while step k + 1 runs): world-size-1 NCCL group,
    the gathered copy of every step must be that step's records even though the source buffer is overwritten right after."""
    import os
    import socket
    import torch.distributed as dist
    from ofdm_sync_math_b200 import dist as odist
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
    try:
        rec = torch.zeros((1, 4096), dtype=torch.uint8, device="cuda")
        g = odist.PipelinedGatherer(rec)
        for k in range(7):
            rec.fill_(k + 1)
            g.push()
            rec.fill_(255)                                   # "the next step" scribbles over the live buffer
            if k >= 1:
                g.drain()
                torch.cuda.synchronize()
                assert int(g.result(k)[0][0, 0]) == k + 1 and int(g.result(k)[0].min()) == k + 1
                assert int(g.result(k - 1)[0][0, 17]) == k
    finally:
        dist.destroy_process_group()
