"""ofs_sync_host pipeline probe: e2e Msamples/s of bench.py's e2e leg (1024 frames x 262144 c64 from pinned host memory,
M + records back to host) for several pipeline batch sizes (OFS_HOST_BATCH_MB), plus plain H2D / D2H copy rates for scale."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ofdm_sync_math_b200 import engine, synth

F, n = 1024, 262144
dev = torch.device("cuda", 0)
x = synth.make_batch_device(F, n, "sc", seed=1234, device=dev)
xh = torch.empty((F, n), dtype=torch.complex64).pin_memory(); xh.copy_(x)
Mh = torch.empty((F, n - 2047), dtype=torch.float32).pin_memory()
rh = torch.zeros((F, engine.REC_BYTES), dtype=torch.uint8).pin_memory()
xd = torch.empty_like(x)
torch.cuda.synchronize()
for name, fn, nbytes in (("h2d only", lambda: xd.copy_(xh, non_blocking=True), xh.numel() * 8),
                         ("d2h only", lambda: xh.copy_(xd, non_blocking=True), xh.numel() * 8)):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    print(json.dumps({"case": name, "GBps": 3 * nbytes / (time.perf_counter() - t0) / 1e9}), flush=True)
xh.copy_(x); torch.cuda.synchronize()
kw = dict(kind="sc", symbol_len=2048, cp_len=512, smooth_win=16, sc_delta=16)
for mb in (256, 128, 64, 32, 16, 8):
    os.environ["OFS_HOST_BATCH_MB"] = str(mb)
    hs = engine.HostSync(0)
    for with_m in (True, False):
        hs.run(xh, Mh if with_m else None, rh, **kw); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(4): hs.run(xh, Mh if with_m else None, rh, **kw)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        print(json.dumps({"case": "ofs_sync_host", "batch_mb": mb, "M_to_host": with_m, "ms": dt * 1e3, "Msamples_per_s": F * n / dt / 1e6,
                          "h2d_GBps": F * n * 8 / dt / 1e9}), flush=True)
    hs.close()
