"""cfg 4 zc_freq metric in FFT form for the profiler: python profiles/prof_zcfreq.py [captures] (ofs_zc_freq_metric_fft, 3 runs)."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ofdm_sync_math_b200 import engine, synth
from ofdm_sync_math_b200.zc import generate_zadoff_chu
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
x = synth.make_batch_device(F, 65536, "sc", seed=11, chunk=64)[:, None]
bi = np.concatenate((np.arange(-31, 0), np.arange(1, 32)))
tb = generate_zadoff_chu(25, 62)
for _ in range(3):
    m = engine.zc_freq_metric(x, bi, tb, 62.0, out_f64=False, fast="fft")
torch.cuda.synchronize()
print("max", float(m.max()))
