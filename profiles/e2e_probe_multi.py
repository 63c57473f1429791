"""Concurrent host<->device ceiling of the box, N ranks at once (the N = 8 analogue of e2e_probe.py):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 profiles/e2e_probe_multi.py
Every rank owns one GPU and 2 GB of pinned host memory.  Measured with all ranks running at the same time (barrier before, max
over ranks): plain pinned H2D, D2H, both directions at once, and bench.py's e2e leg (ofs_sync_host: x in, M + records out).
Rank 0 prints one JSON object: per-rank and aggregate GB/s, the box's CPU / NUMA layout, and e2e as a fraction of the
concurrent H2D ceiling."""
import json
import os
import subprocess
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ofdm_sync_math_b200 import engine, synth  # noqa: E402

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
F, n = 1024, 262144
x = synth.make_batch_device(F, n, "sc", seed=1234 + rank, device=dev)
xh = torch.empty((F, n), dtype=torch.complex64).pin_memory(); xh.copy_(x)
Mh = torch.empty((F, n - 2047), dtype=torch.float32).pin_memory()
Md = torch.empty((F, n - 2047), dtype=torch.float32, device=dev)
rh = torch.zeros((F, engine.REC_BYTES), dtype=torch.uint8).pin_memory()
xd = torch.empty_like(x)
s2 = torch.cuda.Stream()


def sync_all():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def timed(fn, reps=3):
    fn(); sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / reps], device=dev, dtype=torch.float64)
    mine = float(dt.item())
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    return mine, float(dt.item())


def both():
    xd.copy_(xh, non_blocking=True)
    with torch.cuda.stream(s2):
        Mh.copy_(Md, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


out = {"n_gpus": world, "bytes_h2d": xh.numel() * 8, "bytes_d2h": Mh.numel() * 4}
for name, fn, nb in (("h2d", lambda: xd.copy_(xh, non_blocking=True), xh.numel() * 8),
                     ("d2h", lambda: Mh.copy_(Md, non_blocking=True), Mh.numel() * 4),
                     ("h2d+d2h", both, xh.numel() * 8 + Mh.numel() * 4)):
    mine, worst = timed(fn)
    out[name] = {"GBps_per_rank_slowest": nb / worst / 1e9, "GBps_aggregate": world * nb / worst / 1e9, "GBps_this_rank": nb / mine / 1e9}
kw = dict(kind="sc", symbol_len=2048, cp_len=512, smooth_win=16, sc_delta=16)
hs = engine.HostSync(local)
for with_m in (True, False):
    mine, worst = timed(lambda: hs.run(xh, Mh if with_m else None, rh, **kw), reps=4)
    out["ofs_sync_host_M" if with_m else "ofs_sync_host_records_only"] = {
        "Msamples_per_s_aggregate": world * F * n / worst / 1e6, "h2d_GBps_aggregate": world * F * n * 8 / worst / 1e9,
        "fraction_of_concurrent_h2d_ceiling": (world * F * n * 8 / worst / 1e9) / out["h2d"]["GBps_aggregate"],
        "fraction_of_concurrent_bidirectional_ceiling": (world * (F * n * 8 + (Mh.numel() * 4 if with_m else 0)) / worst / 1e9) / out["h2d+d2h"]["GBps_aggregate"]}
hs.close()
if rank == 0:
    try:
        out["cpu_count"] = os.cpu_count()
        out["affinity"] = sorted(os.sched_getaffinity(0))[:4] + ["...", len(os.sched_getaffinity(0))]
        out["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout.splitlines()[:12]
        out["numa_nodes"] = sorted(p.name for p in Path("/sys/devices/system/node").glob("node[0-9]*"))
    except Exception as e:
        out["topo_error"] = str(e)
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
