#!/usr/bin/env python
"""bench.py -- BASELINE.json headline: Msamples/s of "sync metric + CFO" on B200, % of the HBM roofline.

Workload (config.workload, BASELINE.json configs[1]): sc.py Schmidl-Cox timing metric + plateau detector + CFO on
4096 synthetic frames x 262 144 complex64 samples per GPU (sc.py frames tiled through cir1/cir2 + AWGN + CFO sweep,
ofdm_sync_math_b200/synth.py).  One step = one pass of the whole batch through
    ofs_metric (stripe kernel)  ->  ofs_sync_detect (plateau finder + P/CFO records)
with the inputs resident in HBM (`value`), and through ofs_sync_host with HOST buffers (`e2e`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--frames F] [--samples S]
For N > 1 launch with torchrun (one rank per GPU); frames are sharded (weak scaling: F frames PER GPU), no
collective on the sample path, one all_gather of the detection records per step (overlapped with the next step).
--impl reference times the CPU restatement of the reference's algorithm (oracle/, kind "port": the reference is
pure Python and cannot travel to the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Msamples/s (complex baseband, whole box) sync metric+CFO"
N_FFT, CP_LEN, SMOOTH, SC_DELTA = 2048, 512, 16, 16


def peaks() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class StdoutToStderr:
    """While active, whatever is written to file descriptor 1 goes to stderr.  NCCL prints its version banner on stdout when
    NCCL_DEBUG is set in the environment; with this around the process-group set-up and the run, rank 0's stdout carries
    exactly one line, the JSON.  Any failure to juggle the descriptors leaves stdout as it was."""

    def __init__(self):
        self.saved = None

    def start(self):
        try:
            sys.stdout.flush()
            self.saved = os.dup(1)
            os.dup2(2, 1)
        except OSError:
            self.saved = None
        return self

    def stop(self):
        if self.saved is None:
            return
        try:
            sys.stdout.flush()
            os.dup2(self.saved, 1)
            os.close(self.saved)
        except OSError:
            pass
        self.saved = None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  The region is K steps of ~2.4 ms -- far shorter than
    nvidia-smi's 100 ms loop -- so a thread polls NVML directly every ~2 ms (nvmlDeviceGetClockInfo +
    nvmlDeviceGetCurrentClocksEventReasons; the recipe's nvidia-smi query reads the same counters).  If the region is too short to
    hold even one sample, one is taken right after its last kernel was queued (the GPU is still executing it).
    Falls back to `nvidia-smi -lms` when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index
        self.nv, self.h, self.stop_flag, self.thread = None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices; honour CUDA_VISIBLE_DEVICES when it lists indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[index]) if index < len(ids) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            try:
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            self.rows.append((time.time(), float(sm), int(rs)))
        except Exception:
            pass

    def _poll(self):
        while not self.stop_flag:
            self._sample()
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark_queued(self):
        """Call right after the last kernel of the timed region was queued (before the synchronize): guarantees a sample
        taken while the region is executing."""
        if self.nv is not None:
            self._sample()

    def stop(self, t0: float, t1: float) -> dict:
        if self.nv is not None:
            self.stop_flag = True
            if self.thread is not None:
                self.thread.join(timeout=1.0)
            nv = self.nv
            try:
                mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            except Exception:
                mx = None
            inside = [(sm, rs) for ts, sm, rs in self.rows if t0 <= ts <= t1]
            reasons = set()
            for _, rs in inside:
                for bit, nm in self.REASONS.items():
                    if rs & bit:
                        reasons.add(nm)
            sm = [v for v, _ in inside]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                    "samples": len(sm), "source": "nvml, polled every ~2 ms inside the timed region"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [c.strip() for c in line.split(",")]
            if len(f) < 8 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for nm, val in zip(names, f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi -lms 100"}


def cpu_port_rate(frames, n_threads: int, results: list | None = None) -> float:
    """Msamples/s of the CPU restatement (oracle/) of sc_streaming_metric + find_plateau_end + CFO on `frames`
    (float64, reference operation order).  results (optional) receives (plateau_end, cfo) per frame."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as orc

    def one(row):
        M, P, R = orc.sc_streaming_metric(row.astype(np.complex128), N_FFT)
        end = orc.find_plateau_end_from_metric(M, CP_LEN, CP_LEN // 4, SMOOTH)
        c = max(end - SC_DELTA, 0)
        return end, -np.angle(P[c]) / (2 * np.pi * (N_FFT // 2))

    orc.lib()
    t0 = time.perf_counter()
    with ThreadPoolExecutor(n_threads) as ex:      # ctypes releases the GIL inside the C oracle
        res = list(ex.map(one, frames))
    dt = time.perf_counter() - t0
    if results is not None:
        results.extend(res)
    return frames.shape[0] * frames.shape[1] / dt / 1e6


def run_reference(args) -> None:
    """--impl reference: the CPU implementation of the path on the host cores (rank 0 only)."""
    import numpy as np
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from ofdm_sync_math_b200 import synth
    cores = os.cpu_count() or 1
    n_frames = max(2 * cores, 16)
    cirs = synth.load_cirs()
    frames = np.stack([synth.tiled_stream_host(args.samples, f, "sc", cirs["cir1" if f % 2 == 0 else "cir2"][1],
                                               [0.0, 5.0, 10.0, 15.0, 20.0][f % 5], -10e3 + 20e3 * f / max(n_frames - 1, 1))
                       for f in range(n_frames)])
    for _ in range(args.warmup):
        cpu_port_rate(frames[:cores], cores)
    rates = [cpu_port_rate(frames, cores) for _ in range(args.steps)]
    v = float(statistics.median(rates))
    sample = f"{n_frames} frames x {args.samples} complex64 per step, {cores} threads, C restatement of sc.py:42-146 (oracle/)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_frames * args.samples / (v * 1e6), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, args.gpus),
        "reference_arm": {"frames_per_step": n_frames, "what": "bounded sample of the workload per step (the CPU needs ~2 s for 32 frames)"},
        "cpu_baseline": {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference itself is pure Python (0.21 Msamples/s/core for sc_streaming_metric, BASELINE.md); "
                "this arm is its C restatement, the fastest honest CPU form of the same algorithm",
    }))


def config_dict(args, world: int) -> dict:
    """The `config` object -- identical in both arms (ours / --impl reference) for the same command line."""
    F, n = args.frames, args.samples
    return {"workload": workload_name(args), "frames_per_gpu": F, "samples_per_frame": n, "n_fft": N_FFT, "cp_len": CP_LEN,
            "smooth_win": SMOOTH, "sc_delta": SC_DELTA,
            "l2": f"inputs larger than L2 ({F * n * 8 / 1e9:.2f} GB of samples per GPU, no flush needed)",
            "e2e_frames_per_step": min(args.e2e_frames, F),
            "parallelism": f"frames sharded over {world} GPU(s), no sample-path collective"}


def workload_name(args) -> str:
    return (f"sc.py Schmidl-Cox timing metric + plateau detector + CFO, {args.frames} frames x {args.samples} complex64 "
            f"per GPU, channel.py-style cir1/cir2 + AWGN + CFO sweep")


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=4096, help="frames per GPU")
    ap.add_argument("--samples", type=int, default=262144)
    ap.add_argument("--e2e-frames", type=int, default=1024, help="frames per GPU pushed through ofs_sync_host per e2e step")
    ap.add_argument("--store-mode", type=int, default=0)
    ap.add_argument("--tma-mode", type=int, default=2, help="2: tiled tensor-map copies with 128B swizzle, 1: 1-D bulk copies")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` object (the other BASELINE.json configs)")
    ap.add_argument("--configs-scale", type=float, default=1.0, help="batch scale of the `configs` measurements")
    ap.add_argument("--no-exact", action="store_true", help="decide on the float32 metric only (no float64 re-evaluation of in-band decisions)")
    ap.add_argument("--exact-band", type=float, default=0.0, help="relative half-width of the float32 uncertainty band (0: library default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from ofdm_sync_math_b200 import _lib, dist as odist, engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    quiet = StdoutToStderr().start() if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    F, n = args.frames, args.samples
    x = synth.make_batch_device(F, n, "sc", seed=1234 + rank, device=dev)
    plan = engine.SyncPlan(F, n, "sc", N_FFT, "c64", cp_len=CP_LEN, smooth_win=SMOOTH, sc_delta=SC_DELTA, store_mode=args.store_mode,
                           tma_mode=args.tma_mode, exact=not args.no_exact, exact_band=args.exact_band)
    # detection records of step k are all-gathered on a side stream while step k + 1 computes (dist.PipelinedGatherer)
    gather = odist.PipelinedGatherer(plan.rec) if world > 1 else None

    def step():
        plan.run_metric_only(x)
        plan.run_detect_only(x)
        if gather is not None:
            gather.push()

    for _ in range(args.warmup):
        step()
    if gather is not None:
        gather.drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()

    # ---- timed region: K steps, CUDA events on the launching stream; the dominant kernel is bracketed separately
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    l0 = _lib.launch_count()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    for k in range(args.steps):
        ev[k][0].record()
        plan.run_metric_only(x)
        ev[k][1].record()
        plan.run_detect_only(x)
        if gather is not None:
            gather.push()
            if k == args.steps - 1:
                gather.drain()                  # every gather finishes inside the timed region
        ev[k][2].record()
    sampler.mark_queued()
    torch.cuda.synchronize()
    t_wall1 = time.time()
    if world > 1:
        dist.barrier()
    launches = _lib.launch_count() - l0
    clocks = sampler.stop(t_wall0, t_wall1)
    total_ms = ev[0][0].elapsed_time(ev[-1][2])
    kern_ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * F * n / (ms_per_step * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (metric_stripe_kernel): algorithmic bytes = 8 B in + 4 B out per sample
    hbm_peak, peak_kind = peaks()
    out_len = n - N_FFT + 1
    alg_bytes = F * (8 * n + 4 * out_len)
    k_ms = float(statistics.mean(kern_ms))
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel, recorded per FRAME so that it
    # scales with the launch it is printed next to (profiles/traffic.json names the frame count of the capture)
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "traffic.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            per_frame = float(tj["metric_stripe_kernel_bytes_per_launch"]) / float(tj.get("frames", 1024))
            if int(tj.get("samples_per_frame", 262144)) == n:
                traffic = per_frame * F
                traffic_src = f"ncu capture at {int(tj.get('frames', 1024))} frames x {n} ({tj.get('source', '')}), scaled to {F} frames"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": "metric_stripe_kernel<4,SC,c64>", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved / hbm_peak, "peak_kind": peak_kind, "traffic": traffic, "traffic_source": traffic_src, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_share_of_step": k_ms / ms_per_step,
                "frac_of_8TBs_spec": achieved / 8000.0}

    # ---- e2e: host buffers through ofs_sync_host (H2D + kernels + D2H inside the timed region)
    Fe = min(args.e2e_frames, F)
    xh = torch.empty((Fe, n), dtype=torch.complex64).pin_memory()
    xh.copy_(x[:Fe])
    Mh = torch.empty((Fe, out_len), dtype=torch.float32).pin_memory()
    rh = torch.zeros((Fe, engine.REC_BYTES), dtype=torch.uint8).pin_memory()
    hs = engine.HostSync(local)
    kw = dict(kind="sc", symbol_len=N_FFT, cp_len=CP_LEN, smooth_win=SMOOTH, sc_delta=SC_DELTA, exact=not args.no_exact,
              exact_band=args.exact_band)
    hs.run(xh, Mh, rh, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e_steps = max(2, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        rec_h = hs.run(xh, Mh, rh, **kw)
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Fe * n * e_steps / float(te.item()) / 1e6
    # the same call without the metric rows going back (M_host = None: records only -- what a caller that only wants timing and
    # CFO does); reported next to the headline figure, which keeps reading M back
    timing_with_m = rec_h["timing"].copy()               # rec_h views the pinned record buffer the next calls overwrite
    rh.zero_()
    hs.run(xh, None, rh, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        rec_ro = hs.run(xh, None, rh, **kw)
    torch.cuda.synchronize()
    tro = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tro, op=dist.ReduceOp.MAX)
    e2e_records_only = world * Fe * n * e_steps / float(tro.item()) / 1e6
    ro_match = bool((rec_ro["timing"] == timing_with_m).all())
    # what the box can move at most: the same pinned buffers through plain copies, all ranks at once (H2D of x alone, and H2D of x
    # with D2H of M on a second stream) -- the ceiling e2e is judged against (profiles/r2_e2e_ceiling_n{2,8}.json: the host side
    # gives 55 GB/s to one GPU alone and 184 GB/s to eight at once)
    xd_probe = x[:Fe]
    s2 = torch.cuda.Stream()

    def copies(bidir: bool) -> float:
        def go():
            xd_probe.copy_(xh, non_blocking=True)
            if bidir:
                with torch.cuda.stream(s2):
                    Mh.copy_(plan.M[:Fe], non_blocking=True)
                torch.cuda.current_stream().wait_stream(s2)
        go(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0c = time.perf_counter()
        for _ in range(2):
            go()
        torch.cuda.synchronize()
        tc = torch.tensor([(time.perf_counter() - t0c) / 2], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tc, op=dist.ReduceOp.MAX)
        return float(tc.item())
    t_h2d, t_bidir = copies(False), copies(True)          # (xh holds a copy of x[:Fe]: the device input is unchanged)
    ceiling = {"h2d_GBps": world * Fe * n * 8 / t_h2d / 1e9, "h2d_plus_d2h_GBps": world * (Fe * n * 8 + Fe * out_len * 4) / t_bidir / 1e9,
               "what": "plain pinned-memory copies of the same buffers, all ranks at once (x in alone; x in + M out on two streams)"}
    e2e_gbps = world * (Fe * n * 8 + Fe * (out_len * 4 + engine.REC_BYTES)) * e_steps / float(te.item()) / 1e9
    # parity of the two paths on the same frames (indices must agree)
    rec_d = plan.records_numpy()
    e2e_match = bool((rec_h["timing"] == rec_d["timing"][:Fe]).all())
    hs.close()
    e2e = {"value": e2e_value, "unit": "Msamples/s", "h2d_bytes_per_step": Fe * n * 8, "d2h_bytes_per_step": Fe * (out_len * 4 + engine.REC_BYTES),
           "frames_per_step": Fe, "steps": e_steps, "records_match_device_path": e2e_match,
           "GBps_both_directions": e2e_gbps, "host_link_ceiling": ceiling,
           "fraction_of_bidirectional_ceiling": e2e_gbps / ceiling["h2d_plus_d2h_GBps"],
           "records_only": {"value": e2e_records_only, "unit": "Msamples/s", "d2h_bytes_per_step": Fe * engine.REC_BYTES,
                            "records_equal": ro_match, "what": "same call with M_host = NULL: only the 32-byte records come back"},
           "api": "ofs_sync_host (pinned host x -> M + records in host memory)"}

    # ---- CPU baseline: the oracle port on a bounded sample of the same workload (rank 0, N = 1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        ns = max(4 * cores, 32)
        sample = x[:ns].cpu().numpy()
        ref = []
        v = cpu_port_rate(sample, cores, ref)
        # parity in the same run: the records of the timed GPU path against the oracle's float64
        # sc_streaming_metric + find_plateau_end_from_metric + CFO on the same frames (widened to complex128)
        from oracle import oracle as orc
        t_ref = np.array([r[0] for r in ref], dtype=np.int64)
        cfo_ref = np.array([r[1] for r in ref])
        n_diff = int((rec_d["timing"][:ns] != t_ref).sum())
        cfo_err = float(np.max(np.abs(rec_d["cfo"][:ns].astype(np.float64) - cfo_ref)) * 2 * np.pi)
        Mo = np.stack([orc.metric_prefix_c64(sample[i], N_FFT, 0) for i in range(2)])
        merr = float(np.max(np.abs(plan.M[:2].cpu().numpy() - Mo) / np.maximum(Mo, 1e-6)))
        cpu = {"value": v, "unit": "Msamples/s", "cores": cores, "kind": "port",
               "sample": f"{ns} frames x {n} complex64 of this workload, {cores} threads, C restatement of sc.py:42-146 (oracle/)",
               "parity_in_run": {"frames_checked": ns, "timing_indices_differing_from_float64_oracle": n_diff,
                                 "timing_indices_equal": n_diff == 0, "max_cfo_err_rad_per_sample": cfo_err,
                                 "max_rel_metric_err": merr,
                                 "exact_reevaluated_frames": int(((rec_d["status"] & 1) != 0).sum()),
                                 "exact_moved_indices": int(((rec_d["status"] & 2) != 0).sum()),
                                 "unresolved_frames": int(((rec_d["status"] & 4) != 0).sum())}}

    # ---- the other BASELINE.json configs (cfg 1, 3, 4, 5), a few timed launches each on this GPU (N = 1 only)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        import bench_configs
        del plan, x, xh, Mh
        torch.cuda.empty_cache()
        try:
            configs = bench_configs.baseline_configs(args.configs_scale, local)
        except Exception as e:           # the headline line must survive a failure here
            configs = {"error": f"{type(e).__name__}: {e}"}

    if quiet is not None:
        quiet.stop()
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": config_dict(args, world),
            "kernel_options": {"store_mode": args.store_mode, "tma_mode": args.tma_mode, "exact": not args.no_exact,
                               "exact_band": args.exact_band or 1e-5},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "configs": configs,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
