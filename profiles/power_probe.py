#!/usr/bin/env python
"""Is the stripe kernel's in-loop time (2.25 ms) a bandwidth or a power limit?  Times the same launch back to back and with
idle gaps between launches, sampling nvidia-smi clocks / power.  (ncu, which isolates every launch, reports 1.72 ms.)"""
import subprocess, sys, time, json
sys.path.insert(0, ".")
import torch
from ofdm_sync_math_b200 import engine, synth
dev = torch.device("cuda", 0)
F, n = 4096, 262144
x = synth.make_batch_device(F, n, "sc", seed=1, device=dev)
plan = engine.SyncPlan(F, n, "sc", 2048, "c64")
def smi():
    o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active", "--format=csv,noheader"],
                       capture_output=True, text=True).stdout.strip()
    return o
for _ in range(3): plan.run_metric_only(x)
torch.cuda.synchronize()
for gap in (0.0, 0.002, 0.02, 0.2):
    ts = []
    for k in range(30 if gap < 0.1 else 8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.run_metric_only(x); e1.record()
        if gap: torch.cuda.synchronize(); time.sleep(gap)
        ts.append((e0, e1))
    torch.cuda.synchronize()
    ms = [a.elapsed_time(b) for a, b in ts]
    print(json.dumps({"gap_s": gap, "first_ms": ms[0], "mean_last_half_ms": sum(ms[len(ms)//2:]) / (len(ms) - len(ms)//2), "min_ms": min(ms), "smi": smi()}), flush=True)
# sustained: 400 launches back to back, sample clocks in the middle
t0 = time.time()
es = [torch.cuda.Event(enable_timing=True) for _ in range(401)]
es[0].record()
for k in range(400):
    plan.run_metric_only(x); es[k + 1].record()
    if k == 200: s_mid = smi()
torch.cuda.synchronize()
ms = [es[k].elapsed_time(es[k + 1]) for k in range(400)]
print(json.dumps({"sustained_400": True, "ms_first10": sum(ms[:10]) / 10, "ms_last100": sum(ms[-100:]) / 100, "smi_mid": s_mid}))
