"""GPU parity, randomised: the gate / hysteresis state machines and the longest-run search against the oracle's serial loops
(sync_aa.py:495-568, zc_v2.py:360-450, minn_rtl.py:750-825, minn.py:131-205) on random masks and metrics.

The kernels build the gates as interval logic over a bitmask, in parallel (prefix-max / suffix-min over mask words, run
summaries joined by a tree); these cases sweep the things that logic depends on: run density, hysteresis shorter and longer
than the gaps, row lengths around the 32-bit word and per-thread slice boundaries, more gates than event slots, batches of
rows.  Everything integer must be equal."""
import numpy as np
import pytest
import torch

from oracle import oracle as orc

pytestmark = pytest.mark.gpu

LENGTHS = [1, 2, 31, 32, 33, 63, 64, 65, 255, 257, 1000, 8191, 8193, 40000]


def _mask(rng, n, style):
    if style == "sparse":
        return rng.random(n) < 0.02
    if style == "dense":
        return rng.random(n) < 0.95
    if style == "half":
        return rng.random(n) < 0.5
    if style == "ones":
        return np.ones(n, bool)
    if style == "zeros":
        return np.zeros(n, bool)
    # bursts: runs and gaps with geometric lengths
    out = np.zeros(n, bool)
    i, on = 0, bool(rng.integers(0, 2))
    while i < n:
        ln = int(rng.geometric(0.05 if on else 0.02))
        out[i:i + ln] = on
        i += ln
        on = not on
    return out


@pytest.mark.parametrize("style", ["sparse", "dense", "half", "bursts", "ones", "zeros"])
def test_rtl_gate_fsm_random_masks(style):
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(abs(hash(style)) % 1000)
    for n in LENGTHS:
        for hyst in (0, 1, 2, 7, 33, 300):
            above = _mask(rng, n, style)
            valid = np.ones(n, bool)
            valid[: int(rng.integers(0, max(1, n // 4)))] = False
            cp = rng.integers(0, 50, n).astype(np.float64)          # many ties: the RTL rule keeps the LAST maximum
            ev_o, seg_o = orc.detect_minn_rtl(dict(corr_positive=cp, above=above, metric_valid=valid), hysteresis=hyst, timing_offset=-3, max_ev=n + 2)
            ev_g = engine.minn_rtl_events(torch.as_tensor(cp), torch.as_tensor(valid), torch.as_tensor(above), hyst, -3)[0]
            seg_g = [(int(e["gate_start"]), int(e["gate_end"])) for e in ev_g]
            assert seg_g == [tuple(s) for s in seg_o.tolist()], (style, n, hyst)
            closed = [e for e in ev_g if e["closed"]]
            assert [(int(e["peak_index"]), int(e["aux"])) for e in closed] == [(int(r[0]), int(r[1])) for r in ev_o], (style, n, hyst)


@pytest.mark.parametrize("style", ["sparse", "half", "bursts", "dense"])
def test_zc_gate_fsm_random_masks_batched(style):
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(7 + len(style))
    for n in (33, 1000, 8193):
        rows = 5
        mag = rng.random((rows, n))
        mag[:, ::7] = 0.5                                             # ties: first maximum wins (zc_v2.py:417)
        above = np.stack([_mask(rng, n, style) for _ in range(rows)])
        valid = np.ones((rows, n), bool); valid[:, :17] = False
        for hyst in (1, 16, 256):
            ev_g, gm = engine.zc_events(torch.as_tensor(mag), torch.as_tensor(valid), torch.as_tensor(above), 62, hyst)
            for r in range(rows):
                st = orc.ZCState(mag[r], np.zeros(n), np.zeros(n), np.zeros(n), above[r], valid[r])
                ev_o, vals_o, gm_o = orc.detect_zc_peaks(st, 62, hyst, max_ev=n + 2)
                got = [(int(e["peak_index"]), int(e["gate_start"]), int(e["gate_end"]), int(e["aux"])) for e in ev_g[r]]
                assert got == [tuple(int(v) for v in row) for row in ev_o.tolist()], (style, n, hyst, r)
                assert np.array_equal(gm[r].cpu().numpy().astype(bool), gm_o), (style, n, hyst, r)
                assert np.array_equal(np.array([e["value"] for e in ev_g[r]]), vals_o)


def test_gate_count_beyond_event_slots():
    """More gates than OFS_MAX_EVENTS (the reference's lists are unbounded): the drop-in call repeats with as many slots as the
    row needs and returns every gate; an explicit max_events is a hard cap that raises instead of cutting the list silently."""
    from ofdm_sync_math_b200 import engine
    n = 64 * 40
    above = np.zeros(n, bool); above[5::20] = True                   # 128 isolated gates with hysteresis 2
    valid = np.ones(n, bool)
    cp = np.arange(n, dtype=np.float64)
    ev_o, seg_o = orc.detect_minn_rtl(dict(corr_positive=cp, above=above, metric_valid=valid), hysteresis=2, timing_offset=0)
    args = (torch.as_tensor(cp), torch.as_tensor(valid), torch.as_tensor(above), 2, 0)
    ev_g = engine.minn_rtl_events(*args)[0]
    assert len(seg_o) == 128 and len(ev_g) == 128
    assert [(int(e["gate_start"]), int(e["gate_end"])) for e in ev_g] == [tuple(s) for s in seg_o.tolist()]
    assert [int(e["peak_index"]) for e in ev_g if e["closed"]] == [int(r[0]) for r in ev_o]
    with pytest.raises(engine.EventOverflow) as ei:
        engine.minn_rtl_events(*args, max_events=64)
    assert ei.value.needed == 128
    # zc_v2 and sync_aa FSMs share the buffers: same behaviour through their entry points
    mag = np.where(above, 1.0, 0.0)
    ev_z, _ = engine.zc_events(torch.as_tensor(mag), torch.as_tensor(valid), torch.as_tensor(above), 62, 2)
    st = orc.ZCState(mag, np.zeros(n), np.zeros(n), np.zeros(n), above, valid)
    ev_zo, _, _ = orc.detect_zc_peaks(st, 62, 2, max_ev=n + 2)
    assert len(ev_zo) > 64 and [int(e["peak_index"]) for e in ev_z[0]] == [int(r[0]) for r in ev_zo]


@pytest.mark.parametrize("style", ["sparse", "half", "bursts", "dense", "ones"])
def test_minn_peak_longest_run_random(style):
    """find_minn_peak's gate = longest run of (smoothed metric >= threshold * peak), earliest on ties (minn.py:155-182)."""
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(3 + len(style))
    for n in (40, 257, 1000, 8193, 70001, 700001):                   # > 576 k samples: the 768-thread form of minn_peak_kernel
        # a two-level metric: the gate mask of the smoothed metric follows `m`, with equal-length runs to exercise the tie rule
        m = _mask(rng, n, style)
        M = np.where(m, 1.0, 0.05) + 1e-3 * rng.random(n)
        for w, thr in ((1, 0.5), (8, 0.5), (16, 0.7)):
            for dt in (np.float64, np.float32):                      # float32 rows take the aligned vector-load window path
                Md = M.astype(dt)
                pk_o, gate_o, _ = orc.find_minn_peak(Md.astype(np.float64), w, thr)
                peak, span, _ = engine.find_minn_peak(torch.as_tensor(Md)[None], w, thr, None, want_ms=True)
                s, e = (int(v) for v in span[0].tolist())
                seg = np.flatnonzero(gate_o)
                assert int(peak[0]) == pk_o, (style, n, w, thr, dt)
                assert (s, e) == (int(seg[0]), int(seg[-1]) + 1), (style, n, w, thr, dt)


@pytest.mark.parametrize("n", [100, 2049, 4096, 8193, 70000])
def test_zc_detect_bitmask_path_equals_three_step_path(n):
    """ofs_zc_detect (threshold kernel -> bitmask -> gate FSM) against ofs_zc_streaming_detection + ofs_zc_events and the oracle."""
    from ofdm_sync_math_b200 import engine
    rng = np.random.default_rng(n)
    rows = 4
    mag = (rng.random((rows, n)) * 0.2).astype(np.float32)
    for r in range(rows):                                           # a few correlation peaks well above the running mean
        for p in rng.integers(0, n, size=max(1, n // 3000)):
            mag[r, p:p + int(rng.integers(1, 40))] += 3.0
    m = torch.as_tensor(mag).cuda()
    for window, hyst in ((64, 16), (2048, 256)):
        ls, v, ab = engine.zc_streaming_detection(m, window, 64, 15, 0.3)
        ev3, _ = engine.zc_events(m, v, ab, 62, hyst, want_gate_mask=False)
        ev2 = engine.zc_detect(m, window, 64, 15, 0.3, 62, hyst)
        for r in range(rows):
            assert ev2[r].tolist() == ev3[r].tolist(), (n, window, hyst, r)
            st = orc.zc_streaming_detection(mag[r].astype(np.float64), window, 64, 15, 0.3)
            ev_o, vals_o, _ = orc.detect_zc_peaks(st, 62, hyst, max_ev=n + 2)
            got = [(int(e["peak_index"]), int(e["gate_start"]), int(e["gate_end"]), int(e["aux"])) for e in ev2[r]]
            assert got == [tuple(int(x) for x in row) for row in ev_o.tolist()], (n, window, hyst, r)
