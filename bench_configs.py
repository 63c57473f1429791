#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs (bench.py stays the headline / driver contract).

One JSON line per case: throughput in Msamples/s (input samples, all branches), the algorithmic bytes or flops of
SURVEY.md 8d, the achieved GB/s or TFLOP/s from CUDA events, and the fraction of the measured peak.
    python bench_configs.py [--scale 1.0] [--cases minn,combined,park,zc,zcfreq,aa64,rtl,tile]
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"])
    return 6650.0


def timeit(fn, steps=5, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


class _Args:
    def __init__(self, scale, cases):
        self.scale, self.cases = scale, cases


BASELINE_CASES = "rtl,minn,combined,park,zc,zcfreq,bank,aa64,mb"


def baseline_configs(scale: float = 1.0, device_index: int = 0) -> dict:
    """The BASELINE.json configs other than the headline, for bench.py's `configs` object: a few timed launches each on this
    GPU, keyed cfg1 / cfg3_* / cfg4_* / cfg5 (entries that carry a `key`)."""
    out = {}
    run_cases(_Args(scale, BASELINE_CASES), lambda d: out.__setitem__(d["key"], {k: v for k, v in d.items() if k != "key"}) if "key" in d else None,
              device_index)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--cases", default="minn,iq16,chan,rx,combined,park,zc,zcfreq,bank,aa64,rtl,mb,dropin,tile")
    a = ap.parse_args()
    run_cases(a, lambda d: print(json.dumps(d), flush=True))


def run_cases(a, sink, device_index: int = 0):
    import numpy as np
    import torch
    from ofdm_sync_math_b200 import engine, synth
    from ofdm_sync_math_b200.zc import build_pss_symbol, generate_zadoff_chu
    dev = torch.device("cuda", device_index)
    hbm = peaks()
    cases = a.cases.split(",")

    def emit(name, ms, samples, alg_bytes=None, flops=None, note="", key=None):
        d = {"case": name, "ms": ms, "Msamples_per_s": samples / (ms * 1e-3) / 1e6, "note": note}
        if key:
            d["key"] = key
        if alg_bytes is not None:
            d["roofline"] = {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": hbm, "unit": "GB/s",
                             "frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm, "algorithmic_bytes": alg_bytes}
        if flops is not None:
            d["flops"] = {"achieved_tflops": flops / (ms * 1e-3) / 1e12, "fp32_fma_peak_tflops": 74.4,
                          "frac": flops / (ms * 1e-3) / 1e12 / 74.4, "algorithmic_flops": flops}
        sink(d)

    if "minn" in cases:
        # cfg 3: Minn metric + find_minn_peak + CFO, 1 M-sample captures x 1024 streams
        F, n = max(int(1024 * a.scale), 8), 1 << 20
        x = synth.make_batch_device(F, n, "minn", seed=7, device=dev, chunk=16)
        plan = engine.SyncPlan(F, n, "minn", 2048, "c64", smooth_win=16, gate_threshold=0.5)
        ms_k = timeit(lambda: plan.run_metric_only(x))
        ms = timeit(lambda: plan.run(x))
        emit("cfg3 minn metric kernel (stripe, D=512)", ms_k, F * n, alg_bytes=F * (8 * n + 4 * (n - 2047)))
        emit("cfg3 minn metric + find_minn_peak (exact mode) + CFO", ms, F * n, alg_bytes=F * (8 * n + 4 * (n - 2047)), key="cfg3_minn",
             note=f"{F} streams x {n} c64; metric kernel alone {ms_k:.3f} ms; roofline figure = the metric kernel's algorithmic bytes over the whole step")
        del x, plan
    if "iq16" in cases:
        # cfg 2 with int16-IQ ingest (4 B/sample in, 4 B/sample out)
        F, n = max(int(4096 * a.scale), 8), 262144
        xi = torch.randint(-2047, 2048, (F, n, 2), dtype=torch.int16, device=dev)
        plan = engine.SyncPlan(F, n, "sc", 2048, "iq16")
        ms_k = timeit(lambda: plan.run_metric_only(xi))
        ms = timeit(lambda: plan.run(xi))
        emit("cfg2 sc metric kernel, int16-IQ input (stripe)", ms_k, F * n, alg_bytes=F * (4 * n + 4 * (n - 2047)))
        emit("cfg2 sc metric + plateau + CFO, int16-IQ input", ms, F * n, alg_bytes=F * (4 * n + 4 * (n - 2047)))
        del xi, plan
    if "chan" in cases:
        # impairment chain, 32 tx rows -> 1024 streams x 262144 (FIR once per row; noise + CFO + ADC per stream)
        S, n, nb = max(int(1024 * a.scale), 8), 262144, 32
        cirs = synth.load_cirs()
        rng = np.random.default_rng(3)
        n_tx = n - (cirs["cir1"].shape[1] - 1)
        base = torch.as_tensor(np.stack([np.tile(synth.frame(rng, "sc"), (n_tx + 9016) // 9017)[:n_tx] for _ in range(nb)]).astype(np.complex64)).to(dev)
        noise = torch.view_as_complex(torch.randn((S, n, 2), device=dev, dtype=torch.float32))
        rows = (np.arange(S) % nb).astype(np.int32)
        snr = np.array([0.0, 5.0, 10.0, 15.0, 20.0])[np.arange(S) % 5]; cfo = np.linspace(-10e3, 10e3, S); fsc = np.full(S, 4.0)
        ms = timeit(lambda: engine.channel_apply(base, cirs["cir1"][1], row_of_stream=rows, unit_noise=noise, snr_db=snr, cfo_hz=cfo,
                                                 fs=30.72e6, full_scale=fsc, want_iq=True), steps=3, warmup=2)
        emit("8f-1 impairment chain: 1100-tap FIR (overlap-save FFT) + AWGN + CFO + 12-bit ADC -> c64 + int16 IQ", ms, S * n,
             alg_bytes=S * n * (8 + 8 + 8 + 4), note="per stream sample: faded row 8 B (L2) + noise 8 B in, c64 8 B + iq 4 B out")
        del base, noise
    if "rx" in cases:
        # 8f-2 / 8f-3: the stages after the detector, one result per frame: CP-correlation CFO estimators and the pilot / data
        # receive chain (CFO correction, 2 x 2048-point FFT, LS estimate, phase-slope timing, equalise, EVM)
        F, n = max(int(4096 * a.scale), 8), 65536
        x = synth.make_batch_device(F, n, "sc", seed=21, device=dev)[:, None]
        starts = np.full(F, 1337 + 2560, dtype=np.int64)                      # CP start of the first pilot symbol of every capture
        for mode in ("plain", "robust", "peak"):
            ms = timeit(lambda: engine.cp_cfo(x, starts, 2048, 512, 30.72e6, mode), steps=5, warmup=2)
            emit(f"8f-2 CP-correlation CFO estimator, mode {mode} (core.py:179-336), one estimate per capture", ms, F * n,
                 note=f"{F} captures; only the ~3 k samples around each CP are read")
        rng = np.random.default_rng(3)
        pu = (rng.choice([-1.0, 1.0], 1200) + 1j * rng.choice([-1.0, 1.0], 1200)) / np.sqrt(2)
        du = (rng.choice([-1.0, 1.0], 1200) + 1j * rng.choice([-1.0, 1.0], 1200)) / np.sqrt(2)
        cfo = np.full(F, 1000.0)
        ms = timeit(lambda: engine.rx_chain(x, starts, cfo, pu, du), steps=5, warmup=2)
        emit("8f-3 receive chain per capture: CFO correction + pilot / data FFT + LS + phase-slope STO + equalise + EVM (float64)", ms, F * n,
             note=f"{F} captures, 2 OFDM symbols each, 1200 used subcarriers; {F / (ms * 1e-3) / 1e6:.2f} M frames/s")
        del x
    if "combined" in cases:
        F, n = max(int(1024 * a.scale), 8), 1 << 19
        x = synth.make_batch_device(F, n, "minn", seed=8, device=dev, chunk=32)[:, None]

        def run():
            m = engine.metric(x, "minn", 2048, want_pr=False, path="stripe", want_chunk_max=True)
            s = engine.metric(x, "sc_both", 2048, want_pr=False, path="stripe", want_chunk_max=True)
            g = engine.sc_gate(s.M, 0.6, chunk_max=s.chunk_max, toff=2047)
            return engine.find_minn_peak_gated(m.M, 16, g)
        ms = timeit(run, steps=3, warmup=2)
        emit("cfg3 combined: Minn + S&C(both halves) metrics + gate + gated peak", ms, F * n, alg_bytes=2 * F * (8 * n + 4 * (n - 2047)),
             note="two metric passes over the same samples (16 B read + 8 B written per sample algorithmic) + byte gate mask")

        def run_fused():
            m = engine.metric(x, "minn", 2048, want_pr=False, path="stripe", want_chunk_max=True)
            s = engine.metric(x, "sc_both", 2048, want_pr=False, path="stripe", want_chunk_max=True)
            return engine.combined_peak(m.M, s.M, s.chunk_max, 2047, 0.6, 16)
        ms = timeit(run_fused, steps=3, warmup=2)
        same = bool(torch.equal(run()[: F], run_fused()[0]))
        emit("cfg3 combined, fused detector (no gate array): Minn + S&C(both halves) metrics + ofs_combined_peak", ms, F * n,
             alg_bytes=2 * F * (8 * n + 4 * (n - 2047)), key="cfg3_combined",
             note=f"{F} streams x {n} c64; peaks equal to the gate + gated-peak path: {same}")
        del x
    if "park" in cases:
        F, n = max(int(64 * a.scale), 2), 1 << 18
        x = synth.make_batch_device(F, n, "sc", seed=9, device=dev, chunk=16)[:, None]
        import os
        os.environ["OFS_PARK_DIRECT"] = "0"
        ms = timeit(lambda: engine.park_metric(x, 2048), steps=3, warmup=2)
        nout = n - 2048
        emit("cfg3 park metric, block-FFT kernel (band-limited self-convolution: 256-point block transforms summed per anti-diagonal, edge triangles direct)",
             ms, F * n, key="cfg3_park",
             note=f"{F} streams x {n} c64, M + P + E out; ~700 flops per output instead of the direct form's 8192 (1024 complex MACs): "
                  f"{F * nout * 1024 * 8 / (ms * 1e-3) / 1e12:.1f} TFLOP/s direct-form equivalent")
        os.environ["OFS_PARK_DIRECT"] = "1"
        msd = timeit(lambda: engine.park_metric(x, 2048), steps=3, warmup=2)
        os.environ["OFS_PARK_DIRECT"] = "0"
        emit("cfg3 park metric, direct O(h) kernel (1024 complex MACs per output, 8x8 register tiles, de-interleaved smem; OFS_PARK_DIRECT=1)", msd, F * n,
             flops=F * nout * 1024 * 8, key="cfg3_park_direct",
             note=f"{F} streams x {n} c64; FMA-bound (SURVEY 7.3-4); flops = 8 per complex MAC; fp32 peak 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4 TFLOP/s")
        del x
    if "zc" in cases:
        # cfg 4 (single root): overlap-save matched filter + zc_v2 streaming detection + gate FSM
        F, n = max(int(2048 * a.scale), 8), 65536
        x = synth.make_batch_device(F, n, "sc", seed=10, device=dev, chunk=64)[:, None]
        ref = build_pss_symbol(include_cp=False)

        def run():
            corr, mag = engine.zc_matched_filter(x, ref, mode=1, out_f64=False)
            ls, v, ab = engine.zc_streaming_detection(mag, 2048, 64, 15, 0.3)
            return engine.zc_events(mag, v, ab, 2048, 256, want_gate_mask=False)
        ms_mf = timeit(lambda: engine.zc_matched_filter(x, ref, mode=1, out_f64=False), steps=3, warmup=2)
        ms_mag = timeit(lambda: engine.zc_matched_filter(x, ref, mode=1, out_f64=False, want_corr=False), steps=5, warmup=2)
        ms = timeit(run, steps=3, warmup=2)
        emit("cfg4 zc matched filter, corr + |corr| out (8192-point overlap-save blocks, fp32)", ms_mf, F * n, alg_bytes=F * (8 * n + 12 * (n + 2047)),
             note="8 B in + 8 B corr + 4 B |corr| out per sample")
        emit("cfg4 zc matched filter, |corr| only (8192-point overlap-save blocks, fp32)", ms_mag, F * n, alg_bytes=F * (8 * n + 4 * (n + 2047)),
             note=f"{F} captures x {n} c64, one root; 8 B in + 4 B |corr| out per sample (SURVEY 8d K4)", key="cfg4_mf")
        emit("cfg4 zc_v2 pipeline, three separate calls with byte flags: matched filter + running-sum threshold + gate FSM", ms, F * n)
        zplan = engine.ZCDetectPlan(F, 1, n, ref)
        ms_p = timeit(lambda: zplan.run(x), steps=5, warmup=2)
        import time as _time
        torch.cuda.synchronize(); t0 = _time.perf_counter(); evl = zplan.events(); t_ev = (_time.perf_counter() - t0) * 1e3
        emit("cfg4 zc_v2.detect_zc_preamble for the batch on the device (ofs_zc_v2_detect: filter -> |corr| -> bitmask -> gate FSM -> event slots)",
             ms_p, F * n, alg_bytes=F * (8 * n + 4 * (n + 2047)), key="cfg4_zc_v2_pipeline",
             note=f"{F} captures x {n} c64; reading the {sum(len(e) for e in evl)} events back into per-capture host lists: {t_ev:.2f} ms more")
        ms_h = timeit(lambda: engine.zc_v2_detect(x, ref), steps=3, warmup=1)
        emit("cfg4 zc_v2.detect_zc_preamble incl. buffer allocation and host event lists (engine.zc_v2_detect)", ms_h, F * n)

        def run_fused():
            corr, mag = engine.zc_matched_filter(x, ref, mode=1, out_f64=False)
            return engine.zc_detect(mag, 2048, 64, 15, 0.3, 2048, 256)
        ms = timeit(run_fused, steps=3, warmup=2)
        a3, a2 = run()[0], run_fused()
        same = all(u.tolist() == v.tolist() for u, v in zip(a3, a2))
        emit("cfg4 zc_v2 pipeline, threshold kernel and gate FSM exchanging a bitmask (ofs_zc_detect)", ms, F * n,
             note=f"events equal to the three-call path: {same}")
        del x
    if "zcfreq" in cases:
        F, n = max(int(256 * a.scale), 4), 65536
        x = synth.make_batch_device(F, n, "sc", seed=11, device=dev, chunk=64)[:, None]
        half = 31
        bi = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
        tb = generate_zadoff_chu(25, 62)
        ms = timeit(lambda: engine.zc_freq_metric(x, bi, tb, 62.0, out_f64=False), steps=3, warmup=2)
        emit("cfg4 zc_freq metric (62-bin sliding DFT, float64 prefix)", ms, F * n, flops=F * (n - 2559) * 62 * 2 * 16,
             note=f"{F} captures x {n}; flops ~ 62 bins x 2 (tile halo) x ~16 per modulated-prefix sample (float64: the fp32 peak does not apply)",
             key="cfg4_zcfreq_f64")
        Ff = max(int(2048 * a.scale), 4)
        xf = synth.make_batch_device(Ff, n, "sc", seed=11, device=dev, chunk=64)[:, None]
        ms32 = timeit(lambda: engine.zc_freq_metric(xf, bi, tb, 62.0, out_f64=False, fast="f32"), steps=3, warmup=2)
        emit("cfg4 zc_freq metric, float32 sliding-DFT kernel (packed fp32 recurrence, 1e-4 tolerance)", ms32, Ff * n,
             flops=Ff * (n - 2559) * 62 * 24.0, key="cfg4_zcfreq_f32",
             note=f"{Ff} captures x {n}; flops = 62 bins x 12 FMA-class operations per offset (rotation 4, energy 2, template product 4, feed 2)")
        msf = timeit(lambda: engine.zc_freq_metric(xf, bi, tb, 62.0, out_f64=False, fast="fft"), steps=3, warmup=2)
        # 1 forward + 2 inverse 8192-point FFTs per 6145 offsets, 5 N log2 N flops each, + 2 pointwise products
        emit("cfg4 zc_freq metric, FFT form (two overlap-save filters + energy recurrence, float32, 1e-4 tolerance)", msf, Ff * n,
             alg_bytes=Ff * (8 * n + 4 * (n - 2559)), key="cfg4_zcfreq_fft",
             note=f"{Ff} captures x {n}; 8 B in + 4 B metric out per sample; {Ff * ((n - 2559 + 6144) // 6145) * (3 * 5 * 8192 * 13 + 2 * 6 * 8192) / msf / 1e9:.1f} TFLOP/s of FFT arithmetic")
        x2 = xf.reshape(Ff // 2, 2, n)
        ms2 = timeit(lambda: engine.zc_freq_metric(x2, bi, tb, 62.0, out_f64=False, fast="fft"), steps=3, warmup=2)
        emit("cfg4 zc_freq metric, FFT form, 2 receive branches summed (what zc_freq.py's own run feeds)", ms2, Ff * n,
             alg_bytes=Ff // 2 * (16 * n + 4 * (n - 2559)), key="cfg4_zcfreq_fft_2br", note=f"{Ff // 2} captures x 2 branches x {n}")
        del xf, x2
        ms = timeit(lambda: engine.zc_freq_metric(x, bi, tb, 62.0, out_f64=False, fast=True), steps=3, warmup=2)
        emit("cfg4 zc_freq metric, fast path (bank kernel: sliding-DFT recurrence + tcgen05, one template)", ms, F * n,
             alg_bytes=F * (8 * n + 4 * (n - 2559)), note="8 B in + 4 B metric out per sample; FP16 operands, 5e-3 tolerance",
             key="cfg4_zcfreq_fast")
        del x
    if "bank" in cases:
        # cfg 4 (bank): 64 ZC roots x 2048 captures x every offset; fused sliding-DFT producer + tcgen05 kind::f16 MMA
        # (M=128 offsets, N=128 = 64 roots x {Re,Im}, K=128 = 62 bins x {Re,Im} padded)
        F, n = max(int(2048 * a.scale), 2), 65536
        x = synth.make_batch_device(F, n, "sc", seed=14, device=dev, chunk=64)
        half = 31
        bi = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
        T = np.stack([generate_zadoff_chu(r, 62) for r in range(1, 65)])
        ms = timeit(lambda: engine.zc_bank(x, bi, T), steps=3, warmup=2)
        noff = n - 2559
        d = {"case": "cfg4 zc 64-root correlator bank (fused sliding-DFT producers + tcgen05 f16 MMA, fp32 TMEM accumulators + epilogue)", "ms": ms,
             "Msamples_per_s": F * n / (ms * 1e-3) / 1e6, "captures": F,
             "tensor": {"issued_tflops": F * noff * 2.0 * 128 * 128 / (ms * 1e-3) / 1e12,
                        "useful_tflops_8KR": F * noff * 8.0 * 62 * 64 / (ms * 1e-3) / 1e12,
                        "note": "issued = one 128x128x128 fp16 MMA group (fp32 accumulate) per 128 offsets; useful = 8*K*R, K=62 bins, R=64 roots (SURVEY 8d)"}}
        d["key"] = "cfg4_bank"
        sink(d)
        del x
    if "aa64" in cases:
        # cfg 5: 64-antenna [A][A] combining, 8 captures per GPU x 64 antennas x 262144 c64
        F, A, n = max(int(8 * a.scale), 1), 64, 262144
        x = synth.make_batch_device(F * A, n, "sc", seed=12, device=dev, chunk=64).reshape(F, A, n)

        ms_t = timeit(lambda: engine.metric(x, "aa", 512, want_pr=True, out_f64=False, path="tile"), steps=3, warmup=2)
        ms_k = timeit(lambda: engine.metric(x, "aa", 512, want_pr=True, out_f64=False, path="array"), steps=5, warmup=3)
        plan = engine.AADetectPlan(F, A, n, 512, 0.15, 128, 15.36e6)
        ms = timeit(lambda: plan.run(x), steps=5, warmup=3)
        emit("cfg5 sync_aa 64-antenna metric, tile kernel (float64 prefix; M,P,R out)", ms_t, F * A * n, alg_bytes=F * n * (8 * A + 16))
        emit("cfg5 sync_aa 64-antenna metric, array kernel (TMA ring, antenna sum on chip; M,P,R out)", ms_k, F * A * n,
             alg_bytes=F * n * (8 * A + 16), note="8*A B in + M,P,R out per output sample")
        emit("cfg5 sync_aa 64-antenna fused detector: array kernel (M,P + bitmask) + gate FSM + CFO", ms, F * A * n,
             alg_bytes=F * n * (8 * A + 12), key="cfg5", note=f"{F} captures x {A} antennas x {n} c64 per GPU")
        del x, plan
        xi = torch.randint(-2047, 2048, (F, A, n, 2), dtype=torch.int16, device=dev)
        plan = engine.AADetectPlan(F, A, n, 512, 0.15, 128, 15.36e6, in_dtype="iq16")
        ms = timeit(lambda: plan.run(xi), steps=5, warmup=3)
        emit("cfg5 sync_aa 64-antenna fused detector, int16-IQ input", ms, F * A * n, alg_bytes=F * n * (4 * A + 12))
        del xi, plan
    if "rtl" in cases:
        F, A, n = max(int(2048 * a.scale), 16), 2, 32768
        iq = torch.randint(-2047, 2048, (F, A, n, 2), dtype=torch.int16, device=dev)

        def run():
            d = engine.minn_rtl_int(iq, 512, 3, 3276, 15)
            return engine.minn_rtl_events(d["corr_positive"], d["metric_valid"], d["above_threshold"], 2, 0)
        ms_alloc = timeit(run, steps=3, warmup=2)
        rplan = engine.RtlIntPlan(F, A, n, 512, 3, 3276, 15, hysteresis=2, timing_offset=0, max_events=2048)
        ms = timeit(lambda: rplan.run(iq), steps=5, warmup=2)
        emit("cfg1 minn_rtl integer datapath (window sums + combine, parallel floor-shift smoother, threshold) + gate FSM, on the device",
             ms, F * A * n, alg_bytes=F * n * (4 * A + 4 * 8 + 2), key="cfg1_rtl_int",
             note=f"{F} streams x {A} antennas x {n} int16 IQ; all six int64 / uint8 state arrays written (34 B per sample); with "
                  f"allocation of the arrays and host event lists per call (engine.minn_rtl_int + minn_rtl_events): {ms_alloc:.2f} ms")
        del rplan
        del iq
    if "mb" in cases:
        # two branches per frame (what minn.py:347-351 feeds): multi-branch stripe kernel vs the precise tile kernel
        B, n = 2, 262144
        F = max(int(1024 * a.scale), 8)
        x = synth.make_batch_device(F * B, n, "sc", seed=15, device=dev, chunk=64).reshape(F, B, n)
        for kind in ("sc", "minn"):
            ms = timeit(lambda: engine.metric(x, kind, 2048, want_pr=False, path="auto", want_chunk_max=True), steps=5, warmup=3)
            emit(f"{kind} metric, {B} branches summed on chip (multi-branch stripe kernel)", ms, F * B * n, alg_bytes=F * (8 * B * n + 4 * (n - 2047)),
                 key=f"two_branch_{kind}", note=f"{F} frames x {B} branches x {n} c64")
        del x
    if "agree" in cases:
        # how often does the float32 fast path pick a different timing index than a float64 metric on the same input?
        F, n = max(int(1024 * a.scale), 8), 262144
        x = synth.make_batch_device(F, n, "sc", seed=1234, device=dev)
        plan = engine.SyncPlan(F, n, "sc", 2048, "c64")
        fast = plan.run(x).records_numpy()["timing"].copy()
        diff, big = 0, 0
        for f0 in range(0, F, 128):
            r = engine.metric(x[f0:f0 + 128, None], "sc", 2048, want_pr=False, out_f64=True, path="tile")
            t64 = engine.find_plateau_end(r.M, 512, 128, 16).cpu().numpy()
            d = np.abs(t64 - fast[f0:f0 + 128])
            diff += int((d != 0).sum()); big += int((d > 1).sum())
        sink({"case": "agreement of timing indices: float32 stripe path vs float64 tile path (same complex64 input)",
              "frames": F, "frames_with_different_index": diff, "of_which_differ_by_more_than_1": big})
        del x, plan
    if "dropin" in cases:
        # the reference's own call signature: numpy complex128 in -> numpy float64 / complex128 out (H2D + kernel + D2H per call)
        import time
        from ofdm_sync_math_b200 import sc as sc_shim, minn as minn_shim, sync_aa as aa_shim, minn_rtl as rtl_shim, park as park_shim
        rng = np.random.default_rng(0)
        rx = (rng.standard_normal(1 << 20) + 1j * rng.standard_normal(1 << 20)).astype(np.complex128)
        for name, fn in (("sc.sc_streaming_metric", lambda: sc_shim.sc_streaming_metric(rx)),
                         ("minn.minn_streaming_metric", lambda: minn_shim.minn_streaming_metric(rx)),
                         ("sync_aa.aa_detect_streaming", lambda: aa_shim.aa_detect_streaming(rx[: 1 << 18], L=512)),
                         ("minn_rtl.minn_rtl_streaming_metric", lambda: rtl_shim.minn_rtl_streaming_metric(
                             rx[: 1 << 18].reshape(2, -1), smooth_shift=3, threshold_value=3276, threshold_frac_bits=15)),
                         ("park.park_streaming_metric", lambda: park_shim.park_streaming_metric(rx[: 1 << 18]))):
            fn(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            ns = rx.size if name.startswith(("sc.", "minn.")) else 1 << 18
            sink({"case": f"drop-in {name}: numpy complex128 in -> numpy out, one capture of {ns} samples", "ms": dt * 1e3,
                  "Msamples_per_s": ns / dt / 1e6,
                  "note": "wall clock incl. H2D, float64 kernel, D2H; the reference Python runs these at 0.21 / 0.04 / 0.12 Msamples/s (BASELINE.md)"})
    if "tile" in cases:
        F, n = max(int(512 * a.scale), 8), 262144
        x = synth.make_batch_device(F, n, "sc", seed=13, device=dev, chunk=64)[:, None]
        ms = timeit(lambda: engine.metric(x, "sc", 2048, want_pr=True, out_f64=False, path="tile"), steps=3, warmup=2)
        emit("sc metric, precise tile kernel (float64 prefix; M,P,R out)", ms, F * n, alg_bytes=F * (8 * n + 16 * (n - 2047)))
        ms = timeit(lambda: engine.metric(x, "sc", 2048, want_pr=True, out_f64=False, path="stripe"), steps=3, warmup=2)
        emit("sc metric, stripe kernel with P and R outputs (M,P,R out)", ms, F * n, alg_bytes=F * (8 * n + 16 * (n - 2047)))
        del x


if __name__ == "__main__":
    main()
