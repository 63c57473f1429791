"""cfg 1 integer pipeline for the profiler: python profiles/prof_rtl.py [frames] (ofs_minn_rtl_int + gate FSM, 3 runs)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ofdm_sync_math_b200 import engine
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
iq = torch.randint(-2047, 2048, (F, 2, 32768, 2), dtype=torch.int16, device="cuda")
for _ in range(3):
    d = engine.minn_rtl_int(iq, 512, 3, 3276, 15)
    ev = engine.minn_rtl_events(d["corr_positive"], d["metric_valid"], d["above_threshold"], 2, 0)
torch.cuda.synchronize()
print("events", sum(len(e) for e in ev))
