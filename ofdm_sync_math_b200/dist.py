"""Multi-GPU plumbing: one process per GPU (torchrun), captures sharded over ranks, NO collective on the
sample path; the only exchange is an all-gather of the fixed-size detection records.
Works with backend "nccl" (GPU tensors over NVLink) and "gloo" (CPU tensors; used by the CPU tests)."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def init(backend: str | None = None, timeout_s: float = 120.0, device_id=None) -> None:
    """init_process_group with a bounded collective timeout (a rank that died must not hang the others for
    the default 10-30 minutes): the records all_gather is the only collective of the engine and moves < 1 MB, so two minutes
    mean a lost peer, not a slow one.  RANK / WORLD_SIZE / MASTER_* come from the environment (torchrun)."""
    from datetime import timedelta
    if dist.is_initialized():
        return
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    kw = {"timeout": timedelta(seconds=float(timeout_s))}
    if device_id is not None and backend == "nccl":
        kw["device_id"] = device_id
    dist.init_process_group(backend, **kw)


def shard_range(n_items: int, rank: int | None = None, world: int | None = None) -> tuple[int, int]:
    """Contiguous block [lo, hi) of the batch axis owned by `rank` (sizes differ by at most one)."""
    world = dist.get_world_size() if world is None else world
    rank = dist.get_rank() if rank is None else rank
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_sizes(n_items: int, world: int) -> list[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


class RecordGatherer:
    """All-gather of a per-rank records tensor [frames_local, record_bytes] (uint8) into [world, frames_max, record_bytes].
    Buffers are allocated once; ranks with fewer frames are zero-padded (counts travel with the records)."""

    def __init__(self, rec: torch.Tensor, frames_max: int | None = None):
        self.world = dist.get_world_size()
        self.rec = rec
        self.frames_max = int(frames_max if frames_max is not None else rec.shape[0])
        self.padded = rec if rec.shape[0] == self.frames_max else torch.zeros((self.frames_max, rec.shape[1]), dtype=rec.dtype, device=rec.device)
        self.out = [torch.zeros_like(self.padded) for _ in range(self.world)]

    def run(self) -> list[torch.Tensor]:
        if self.padded is not self.rec:
            self.padded[: self.rec.shape[0]].copy_(self.rec)
        dist.all_gather(self.out, self.padded)
        return self.out


class PipelinedGatherer:
    """RecordGatherer off the critical path: the records of step k are snapshotted into one of two staging buffers on the
    compute stream (a few KB, device to device) and all-gathered on a side stream while step k + 1 computes.  push() after
    every step; drain() before reading results (or the clock).  result(k) = the gathered list for the step pushed k-th
    (only the two most recent are kept)."""

    def __init__(self, rec: torch.Tensor):
        self.world = dist.get_world_size()
        self.rec = rec
        self.stage = [torch.zeros_like(rec) for _ in range(2)]
        self.out = [[torch.zeros_like(rec) for _ in range(self.world)] for _ in range(2)]
        self.comm = torch.cuda.Stream(device=rec.device)
        self.ready = [torch.cuda.Event() for _ in range(2)]      # snapshot k written (compute stream)
        self.done = [torch.cuda.Event() for _ in range(2)]       # gather k finished (side stream)
        self.k = 0

    def push(self) -> None:
        b = self.k & 1
        cur = torch.cuda.current_stream(self.rec.device)
        if self.k >= 2:
            cur.wait_event(self.done[b])                         # the gather that last read this staging buffer
        self.stage[b].copy_(self.rec, non_blocking=True)
        self.ready[b].record(cur)
        with torch.cuda.stream(self.comm):
            self.comm.wait_event(self.ready[b])
            dist.all_gather(self.out[b], self.stage[b])
            self.done[b].record(self.comm)
        self.k += 1

    def drain(self) -> None:
        cur = torch.cuda.current_stream(self.rec.device)
        for b in range(min(self.k, 2)):
            cur.wait_event(self.done[b])

    def result(self, k: int) -> list[torch.Tensor]:
        return self.out[k & 1]


def gather_records(rec: torch.Tensor, n_total: int) -> np.ndarray | None:
    """Gather every rank's records (sharded with shard_range over n_total frames) and return them in global
    frame order as a uint8 array [n_total, record_bytes] (on every rank)."""
    world = dist.get_world_size()
    sizes = shard_sizes(n_total, world)
    g = RecordGatherer(rec, max(sizes))
    parts = g.run()
    return np.concatenate([parts[r][: sizes[r]].cpu().numpy() for r in range(world)], axis=0)


def gather_records_or_files(rec: torch.Tensor, n_total: int, fallback_dir=None):
    """gather_records, but a failed collective (peer lost, NCCL timeout -- see init()) does not lose this rank's detections:
    they are written to `<fallback_dir>/records_rank<r>_of<w>.npy` (global frame range in the file name) and None is returned,
    so a post-mortem can still assemble the batch from the per-rank files."""
    try:
        return gather_records(rec, n_total)
    except Exception as e:                         # DistBackendError / RuntimeError depending on the backend
        if fallback_dir is None:
            raise
        from pathlib import Path
        d = Path(fallback_dir)
        d.mkdir(parents=True, exist_ok=True)
        rank, world = dist.get_rank(), dist.get_world_size()
        lo, hi = shard_range(n_total, rank, world)
        path = d / f"records_rank{rank}_of{world}_frames{lo}-{hi}.npy"
        np.save(path, rec.detach().cpu().numpy())
        import sys
        print(f"[ofdm_sync_math_b200.dist] records gather failed on rank {rank} ({type(e).__name__}: {e}); wrote {path}", file=sys.stderr)
        return None
