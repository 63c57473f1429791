"""How wide must the exact mode's band be?  Relative error of the stripe kernel's float32 metric against the float64 oracle,
conditional on M / rowmax (decisions of the detectors are taken at >= 0.5 of the row maximum), on bench-recipe frames.
Run on the GPU box: python profiles/exact_band_probe.py > gpurun_out/exact_band_probe.json"""
import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ofdm_sync_math_b200 import engine, synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402

out = {}
for kind, k in (("sc", 0), ("minn", 2)):
    F, n = 96, 262144
    x = synth.make_batch_device(F, n, kind, seed=99)
    M = engine.metric(x[:, None], kind, 2048, want_pr=False, path="stripe").M.cpu().numpy().astype(np.float64)
    xs = x.cpu().numpy()
    worst = {t: 0.0 for t in (0.0, 0.05, 0.1, 0.25, 0.5)}
    worst_sm = dict(worst)
    for f in range(F):
        Mo = orc.metric_prefix_c64(xs[f], 2048, k)
        rel = np.abs(M[f] - Mo) / np.maximum(Mo, 1e-30)
        # 16-sample smoothed values, what the detectors compare
        ker = np.ones(16) / 16
        so, sg = np.convolve(Mo, ker, "same"), np.convolve(M[f], ker, "same")
        rels = np.abs(sg - so) / np.maximum(so, 1e-30)
        for t in worst:
            sel = Mo >= t * Mo.max() if t > 0 else Mo >= 1e-6
            worst[t] = max(worst[t], float(rel[sel].max()))
            sels = so >= t * so.max() if t > 0 else so >= 1e-6
            worst_sm[t] = max(worst_sm[t], float(rels[sels].max()))
    out[kind] = {"frames": F, "max_rel_err_given_M_ge_frac_of_rowmax": worst, "same_for_16_sample_smoothed": worst_sm}
print(json.dumps(out))
