"""N > 1 host logic on CPU: world_size-2 gloo process group, contiguous frame sharding, record all-gather."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, n_total, q):
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ofdm_sync_math_b200 import dist as odist
    lo, hi = odist.shard_range(n_total)
    rec = torch.zeros((hi - lo, 32), dtype=torch.uint8)
    for i in range(hi - lo):                      # record = global frame index, byte-wise
        rec[i] = torch.tensor(np.frombuffer(np.int64(lo + i).tobytes() * 4, dtype=np.uint8))
    allrec = odist.gather_records(rec, n_total)
    ok = allrec.shape == (n_total, 32) and np.array_equal(allrec.view(np.int64)[:, 0], np.arange(n_total))
    q.put((rank, lo, hi, bool(ok)))
    dist.destroy_process_group()


def test_shard_range_partition():
    from ofdm_sync_math_b200 import dist as odist
    for n in (0, 1, 7, 8, 4096, 4097):
        for w in (1, 2, 3, 8):
            r = [odist.shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(w - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
            assert odist.shard_sizes(n, w) == [b - a for a, b in r]


@pytest.mark.parametrize("n_total", [8, 11])
def test_gloo_world2_gather_records(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == n_total
    assert all(r[3] for r in res)


def _worker_events(rank, world, port, caps_per_rank, q):
    """bench_array.py's N > 1 exchange on CPU: every rank owns caps_per_rank captures, fills its event slots + counts (one flat
    buffer, engine._event_buffers(want_flat=True)), all-gathers the flat buffers and decodes everybody's events."""
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ofdm_sync_math_b200 import _lib as L, dist as odist, engine
    ev, cnt, flat = engine._event_buffers(caps_per_rank, torch.device("cpu"), want_flat=True)
    esz = ev.shape[1] // L.OFS_MAX_EVENTS
    evn = ev.numpy().reshape(caps_per_rank, L.OFS_MAX_EVENTS, esz).view(engine._EVENT_NP).reshape(caps_per_rank, L.OFS_MAX_EVENTS)
    for c in range(caps_per_rank):
        k = (rank * caps_per_rank + c) % 5                       # 0..4 events in this capture
        cnt[c] = k
        for e in range(k):
            evn[c, e]["peak_index"] = 1000 * (rank * caps_per_rank + c) + e
            evn[c, e]["cfo"] = 0.5 * e + rank
            evn[c, e]["closed"] = 1
    parts = odist.RecordGatherer(flat.view(1, -1)).run()
    ok = len(parts) == world
    row_b = L.OFS_MAX_EVENTS * esz
    for r in range(world):
        f = parts[r].reshape(-1)
        ev_r = f[: caps_per_rank * row_b].view(caps_per_rank, row_b)
        cnt_r = f[caps_per_rank * row_b:].view(torch.int32)
        dec = engine._events_to_numpy(ev_r, cnt_r)
        for c in range(caps_per_rank):
            g = r * caps_per_rank + c
            ok &= len(dec[c]) == g % 5
            ok &= [int(x) for x in dec[c]["peak_index"]] == [1000 * g + e for e in range(g % 5)]
            ok &= all(float(x) == 0.5 * e + r for e, x in enumerate(dec[c]["cfo"]))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_gloo_world2_gather_event_records():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_events, args=(r, 2, port, 3, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def _worker_fallback(rank, world, port, tmp, q):
    """A failing collective: every rank keeps its records in a per-rank file instead of losing them."""
    sys.path.insert(0, str(ROOT))
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    os.environ["RANK"] = str(rank); os.environ["WORLD_SIZE"] = str(world)
    from ofdm_sync_math_b200 import dist as odist
    odist.init("gloo", timeout_s=30.0)
    lo, hi = odist.shard_range(10)
    rec = torch.full((hi - lo, 40), rank + 1, dtype=torch.uint8)
    ok_path = odist.gather_records_or_files(rec, 10, tmp) is not None       # healthy group: the gather works
    real = dist.all_gather
    dist.all_gather = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("simulated NCCL timeout"))
    try:
        none = odist.gather_records_or_files(rec, 10, tmp)
    finally:
        dist.all_gather = real
    q.put((rank, ok_path, none is None))
    dist.destroy_process_group()


def test_gloo_world2_gather_failure_writes_rank_files(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_fallback, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True)]
    files = sorted(f.name for f in tmp_path.iterdir())
    assert files == ["records_rank0_of2_frames0-5.npy", "records_rank1_of2_frames5-10.npy"]
    assert np.load(tmp_path / files[1]).shape == (5, 40) and int(np.load(tmp_path / files[1])[0, 0]) == 2
