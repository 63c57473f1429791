"""cfg 3 Park metric for the profiler: python profiles/prof_park.py [streams] (3 runs of ofs_park_metric, block-FFT kernel unless OFS_PARK_DIRECT=1)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ofdm_sync_math_b200 import engine, synth
F = int(sys.argv[1]) if len(sys.argv) > 1 else 64
x = synth.make_batch_device(F, 262144, "sc", seed=12, chunk=64)[:, None]
for _ in range(3):
    M, P, E = engine.park_metric(x, 2048)
torch.cuda.synchronize()
print("max", float(M.max()))
