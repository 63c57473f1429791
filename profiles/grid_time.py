import time, json, sys
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from ofdm_sync_math_b200 import sync_aa
kw = dict(snr_values=[-5, 0, 5, 10, 15], channels=[None, "cir1", "cir2"], full_scale_ratios=[0.5, 1.0, 2.0],
          preamble_lengths=[1024, 512, 256], cfo_hz=500.0, plot_samples=False)
for i in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = sync_aa.run_grid_test(**kw)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(json.dumps({"run": i, "cases": len(r), "detected": sum(x.detected for x in r), "seconds": dt}))
