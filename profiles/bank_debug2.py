import numpy as np, sys, torch
sys.path.insert(0,'.')
from ofdm_sync_math_b200 import engine
g=dict(np.load('tests/golden/zc_freq_awgn.npz'))
rx=g["rx"][0].astype(np.complex64)
rng=np.random.default_rng(3)
long_cap=np.concatenate([0.3*(rng.standard_normal(40000)+1j*rng.standard_normal(40000)).astype(np.complex64), rx, 0.3*(rng.standard_normal(30000)+1j*rng.standard_normal(30000)).astype(np.complex64)])
def zc(root,length=62):
    n=np.arange(length); return np.exp(-1j*np.pi*root*n*(n+1)/length)
T=np.stack([zc(25), zc(3)])
bm,bo=engine.zc_bank(long_cap[None], g["bin_indices"], T)
print("bank", bm.cpu().numpy(), bo.cpu().numpy())
m=engine.zc_freq_metric(long_cap[None,None,:], g["bin_indices"], zc(25), 62.0, out_f64=False, fast=True).cpu().numpy()[0]
print("fast argmax", m.argmax(), m.max(), m[41337], m[37517], m.shape)
top=np.argsort(m)[-5:]; print(top, m[top])
m64=engine.zc_freq_metric(long_cap[None,None,:].astype(np.complex128), g["bin_indices"], zc(25), 62.0).cpu().numpy()[0]
print("f64 argmax", m64.argmax(), m64.max())
d=np.abs(m-m64); print("max diff", d.max(), d.argmax())
for n in (20000, 40000, 60000, 82000):
    x=long_cap[:n] if n<=long_cap.size else long_cap
    bm,bo=engine.zc_bank(x[None], g["bin_indices"], T[:1]); print(n, bm.cpu().numpy(), bo.cpu().numpy())
