"""cfg 4 single-root pipeline for the profiler: python profiles/prof_zc.py [captures] (ofs_zc_v2_detect, 3 runs)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ofdm_sync_math_b200 import engine, synth
from ofdm_sync_math_b200.zc import build_pss_symbol
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = synth.make_batch_device(F, 65536, "sc", seed=10, chunk=64)[:, None]
plan = engine.ZCDetectPlan(F, 1, 65536, build_pss_symbol(include_cp=False))
for _ in range(3):
    plan.run(x)
torch.cuda.synchronize()
print("events", sum(len(e) for e in plan.events()))
