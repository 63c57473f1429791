import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
os.environ["OFS_BANK_DEBUG"] = "gpurun_out/bank"
MODE = int(os.environ.get("OFS_BANK_DEBUG_MODE", "0"))
from ofdm_sync_math_b200 import engine
g = dict(np.load("tests/golden/zc_freq_awgn.npz"))
rx = g["rx"][0].astype(np.complex64)[None]
def zc(r, L=62):
    n = np.arange(L); return np.exp(-1j*np.pi*r*n*(n+1)/L)
T = np.stack([zc(r) for r in range(1, 65)])
bm, bo = engine.zc_bank(rx, g["bin_indices"], T)
print("bm[24]", bm[0,24].item(), "bo", bo[0,24].item(), "peak", int(g["peak"]))
n = rx.shape[1]; n_off = n - 2560 + 1; pad = (n_off + 127)//128*128
B = np.fromfile("gpurun_out/bank_binsT.bin", np.float32).reshape(128, pad)
E = np.fromfile("gpurun_out/bank_E.bin", np.float32)
A = np.fromfile("gpurun_out/bank_A.bin", np.float32).reshape(256, 128)
Y = np.fromfile("gpurun_out/bank_Y.bin", np.float32).reshape(256, pad)
print("B nonzero frac", (B != 0).mean(), "E max", E.max(), "A nonzero", (A != 0).mean(), "Y absmax", np.abs(Y).max())
# reference bins via FFT for a few offsets
k = np.mod(g["bin_indices"], 2048)
for o in (0, 5, int(g["peak"])):
    F = np.fft.fft(rx[0, o+512:o+2560].astype(np.complex128))[k]
    print(o, "bins err", np.abs(B[:62, o] + 1j*B[64:126, o] - F).max(), np.abs(F).max(), "E", E[o], np.sum(np.abs(F)**2))
Yref = A @ B     # [256, pad]
print("Y vs A@B max err", np.abs(Y - Yref).max(), np.abs(Yref).max())
# structure of mismatch
d = np.abs(Y - Yref)
print("rows ok:", (d.max(axis=1) < 1e-2*np.abs(Yref).max()).sum(), "cols ok:", (d.max(axis=0) < 1e-2*np.abs(Yref).max()).sum())
np.save("gpurun_out/bank_Ysmall.npy", Y[:, :256]); np.save("gpurun_out/bank_Yref_small.npy", Yref[:, :256])

print("smem A[0:8]", Y[250,:8], "A global row0[0:8]", A[0,:8])
print("smem B[0:8]", Y[251,:8], "B global row0[0:8]", B[0,:8])
print("smem A_im[0:8]", Y[252,:8], "A global row128[0:8]", A[128,:8])
print("tmem base", Y[253,:2].view(np.uint32))

if MODE == 1:
    print("st/ld test: Y[0,:4]", Y[0,:4], "Y[5,:4]", Y[5,:4], "Y[128+5,:4] (cols 128..)", Y[128+5,:4], "Y[37, 100:103]", Y[37,100:103])
if MODE == 2:
    Z = A[:128] @ A[128:].T
    print("A*A^T test: max err", np.abs(Y[:128,:128]-Z).max(), "Zmax", np.abs(Z).max(), "Y sample", Y[0,:4], "Z sample", Z[0,:4])

if MODE >= 3:
    Z = A[:128] @ A[128:]
    print("A x A (MN-major B) test: max err", np.abs(Y[:128,:128]-Z).max(), "Zmax", np.abs(Z).max(), "Y sample", Y[1,:4], "Z sample", Z[1,:4])
    Zt = A[:128] @ A[128:].T
    print("   vs A x A^T err", np.abs(Y[:128,:128]-Zt).max())
