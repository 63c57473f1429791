"""cfg 3 detector probe: Minn metric + find_minn_peak on R rows x 1 M samples (profile target for ncu -k regex:minn_peak)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ofdm_sync_math_b200 import engine, synth
R = int(sys.argv[1]) if len(sys.argv) > 1 else 296
dev = torch.device("cuda", 0)
x = synth.make_batch_device(R, 1 << 20, "minn", seed=7, device=dev, chunk=16)
plan = engine.SyncPlan(R, 1 << 20, "minn", 2048, "c64", smooth_win=16, gate_threshold=0.5)
for _ in range(3):
    plan.run(x)
torch.cuda.synchronize()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); plan.run_metric_only(x); e1.record(); plan.run_detect_only(x); e2.record(); torch.cuda.synchronize()
print("metric ms", e0.elapsed_time(e1), "detect ms", e1.elapsed_time(e2))
