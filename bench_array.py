#!/usr/bin/env python
"""bench_array.py -- BASELINE.json configs[4]: sync_aa.py 64-antenna array combining, captures sharded across the GPUs,
NCCL gather of the detection records only.  (bench.py is the headline, configs[1]; this is the same contract for the
antenna-array workload.)

One step = every capture of this rank through ofs_aa_detect (array kernel: per-antenna P, R over a TMA ring, antenna sum
on chip, M / P rows + threshold bitmask out; then the gate FSM + CFO read-out, sync_aa.py:458-568) followed, for N > 1,
by one all_gather of the event records (64 x 72 B + a count per capture), issued on a side stream so that it overlaps the
next step (dist.PipelinedGatherer; --sync-gather keeps it on the compute stream).  No collective touches the samples.

    python bench_array.py [--gpus N] [--steps K] [--warmup W] [--captures 8] [--antennas 64] [--samples 262144] [--dtype c64|iq16]
For N > 1 launch with torchrun, one rank per GPU (weak scaling: --captures is PER GPU).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from bench import ClockSampler, StdoutToStderr, peaks  # noqa: E402

HALF_LEN, THRESH, HYST, FS = 512, 0.15, 128, 15.36e6


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--captures", type=int, default=8, help="captures per GPU")
    ap.add_argument("--antennas", type=int, default=64)
    ap.add_argument("--samples", type=int, default=262144)
    ap.add_argument("--dtype", default="c64", choices=["c64", "iq16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sync-gather", action="store_true", help="all_gather on the compute stream instead of overlapped with the next step")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from ofdm_sync_math_b200 import _lib, dist as odist, engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    quiet = StdoutToStderr().start() if world > 1 else None      # NCCL's version banner goes to stderr, stdout = the JSON line
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    F, A, n = args.captures, args.antennas, args.samples
    # one host-built capture per rank ([gap][A][A] preambles, per-antenna phase + AWGN, CFO: synth.aa_capture_host), the other
    # captures of the rank are circular shifts of it with their own sign pattern over the antennas (different data, same statistics)
    base = torch.as_tensor(synth.aa_capture_host(n, A, seed=100 + rank, half_len=HALF_LEN, snr_db=10.0, cfo_hz=500.0,
                                                 int12=args.dtype == "iq16")).to(dev)
    g = torch.Generator(device="cpu").manual_seed(7 + rank)
    caps = []
    for f in range(F):
        sh = int(torch.randint(0, n, (1,), generator=g).item()) if f else 0
        caps.append(torch.roll(base, shifts=sh, dims=1))
    x = torch.stack(caps).contiguous()
    del caps, base
    # 64 preambles per capture; a circular shift can cut one in two (65 gates), and the drop-ins no longer cut event lists at
    # the slot count (EventOverflow): 128 slots per capture
    plan = engine.AADetectPlan(F, A, n, HALF_LEN, THRESH, HYST, FS, in_dtype=args.dtype, max_events=128)
    gather = None                                                                      # event slots + counts, one buffer
    if world > 1:
        gather = (odist.RecordGatherer if args.sync_gather else odist.PipelinedGatherer)(plan.records.view(1, -1))

    def step():
        plan.run(x)
        if gather is not None:
            gather.run() if args.sync_gather else gather.push()

    for _ in range(args.warmup):
        step()
    if gather is not None and not args.sync_gather:
        gather.drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = _lib.launch_count()
    torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    if gather is not None and not args.sync_gather:
        gather.drain()                                   # the last gathers finish inside the timed region
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    if world > 1:
        dist.barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop(t0, t1)
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * F * A * n / (ms_per_step * 1e-3) / 1e6
    in_b = 8 if args.dtype == "c64" else 4
    alg = F * n * (in_b * A + 12)                       # every antenna sample once in; M (4 B) + P (8 B) per combined sample out
    hbm, kind = peaks()
    ach = alg / (ms_per_step * 1e-3) / 1e9
    events = plan.events()
    n_ev = int(sum(len(e) for e in events))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.dtype == "c64":
        from oracle import oracle as orc
        ns = min(n, 65536)
        xs = x[0, :, :ns].cpu().numpy().astype(np.complex128)
        tc = time.perf_counter()
        ro = orc.aa_detect_streaming(xs, HALF_LEN, THRESH, HYST, FS)
        dt = time.perf_counter() - tc
        sub = engine.AADetectPlan(1, A, ns, HALF_LEN, THRESH, HYST, FS)
        sub.run(x[:1, :, :ns].contiguous())
        ev_g = sub.events()[0]
        ev_o = ro["ev_i"]                                  # float64 oracle on the complex64 samples: peak, gate start / end, frame start
        same = len(ev_g) == len(ev_o) and all(int(a["peak_index"]) == int(b[0]) for a, b in zip(ev_g, ev_o))
        cpu = {"value": A * ns / dt / 1e6, "unit": "Msamples/s", "cores": 1, "kind": "port",
               "sample": f"1 capture x {A} antennas x {ns} samples, C restatement of sync_aa.py:421-571 (oracle/), one thread",
               "parity_in_run": {"event_peaks_equal": bool(same), "events": len(ev_o)}}

    if quiet is not None:
        quiet.stop()
    if rank == 0:
        print(json.dumps({
            "metric": "Msamples/s (complex baseband antenna samples, whole box) sync_aa array detector", "value": value, "unit": "Msamples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"sync_aa.py {A}-antenna [A][A] array detector, {F} captures x {A} antennas x {n} {args.dtype} per GPU, "
                                   f"captures sharded over {world} GPU(s), NCCL gather of event records only"
                                   + ("" if world == 1 else (" (on the compute stream)" if args.sync_gather else " (overlapped with the next step)")),
                       "captures_per_gpu": F, "antennas": A, "samples": n, "half_len": HALF_LEN,
                       "l2": f"inputs larger than L2 ({x.numel() * x.element_size() / 1e9:.2f} GB per GPU, no flush needed)"},
            "roofline": {"bound": "hbm", "kernel": "aa_array_kernel + aa gate FSM (whole ofs_aa_detect call)", "achieved": ach, "peak": hbm,
                         "unit": "GB/s", "frac": ach / hbm, "peak_kind": kind, "traffic": None, "algorithmic_bytes_per_launch": alg},
            "cpu_baseline": cpu, "events_this_rank": n_ev, "gpu_launches": int(launches), "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
