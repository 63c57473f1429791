"""Small-shape tour of the TMA / mbarrier / tcgen05 kernels for compute-sanitizer (racecheck / memcheck / synccheck):
    compute-sanitizer --tool racecheck python profiles/sanitize_small.py
Shapes are tiny (the tools slow kernels down 10-100x) but take every code path: tiled / 1-D bulk / plain loads of the stripe
kernel, its multi-branch ring, the antenna-array ring, the bank's operand ring + TMEM accumulators, the 8192-point filter,
the threshold kernel and the parallel integer smoother."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ofdm_sync_math_b200 import engine  # noqa: E402
from ofdm_sync_math_b200.zc import build_pss_symbol, generate_zadoff_chu  # noqa: E402

rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)


def cplx(*shape):
    return torch.as_tensor((rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)).to(dev)


# stripe kernel: tiled TMA (pitch % 128 B == 0), 1-D bulk (pitch % 16 B == 0), plain loads (odd pitch); all kinds
for n in (16384, 16386, 16385):
    x = cplx(3, 1, n)
    for kind, N in (("sc", 2048), ("sc_both", 2048), ("minn", 2048), ("aa", 512)):
        engine.metric(x, kind, N, want_pr=False, path="stripe", want_chunk_max=True)
    engine.metric(x, "sc", 2048, want_pr=True, path="stripe")
    engine.metric(x, "sc", 2048, want_pr=False, path="stripe", store_mode=1)
# multi-branch ring
for B in (2, 3):
    engine.metric(cplx(2, B, 16384), "sc", 2048, want_pr=False, path="stripe", want_chunk_max=True)
    engine.metric(cplx(2, B, 16384), "minn", 1024, want_pr=False, path="stripe")
# fused sync with the exact mode
plan = engine.SyncPlan(4, 32768, "sc", 2048, "c64")
plan.run(cplx(4, 32768))
plan = engine.SyncPlan(2, 32768, "minn", 2048, "c64")
plan.run(cplx(2, 32768))
# antenna-array ring + gate FSM
aplan = engine.AADetectPlan(2, 4, 8192, 512, 0.15, 128, 15.36e6)
aplan.run(cplx(2, 4, 8192))
# correlator bank (operand ring, TMEM accumulators)
half = 31
bi = np.concatenate((np.arange(-half, 0), np.arange(1, half + 1)))
T = np.stack([generate_zadoff_chu(r, 62) for r in range(1, 65)])
engine.zc_bank(cplx(2, 12288), bi, T)
# 8192-point matched filter + threshold kernel + FSM
engine.zc_v2_detect(cplx(2, 1, 30000), build_pss_symbol(include_cp=False))
# integer datapath with the parallel smoother
iq = torch.as_tensor(rng.integers(-2047, 2048, size=(2, 2, 20000, 2)).astype(np.int16)).to(dev)
engine.minn_rtl_int(iq, 256, 3, 3276, 15)
torch.cuda.synchronize()
print("sanitize tour done")
