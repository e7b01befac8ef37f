import sys, numpy as np
sys.path.insert(0, '.')
import nnal_b200
eng = nnal_b200.get_engine()
rs = np.random.RandomState(0)
M, N, K = 512, 1024, 4704
A = np.maximum(rs.randn(M, K), 0).astype(np.float32) * 2
W = (rs.randn(N, K) * np.sqrt(2. / K)).astype(np.float32)
b = np.zeros(N, np.float32)
ref = A.astype(np.float64) @ W.astype(np.float64).T
rms = np.sqrt((ref ** 2).mean())
def stats(name, got):
    e = got.astype(np.float64) - ref
    print('%-28s max %.3g  rms %.3g  signed-bias(e*sign(ref)) %.3g   (all / rms(ref)=%.3g)' % (
        name, np.abs(e).max() / rms, np.sqrt((e ** 2).mean()) / rms, (e * np.sign(ref)).mean() / rms, rms))
stats('simt fp32', eng.debug_fc(A, W, b, 0, 0))
stats('tc full K', eng.debug_fc(A, W, b, 0, 1))
for J in (2, 4, 8, 16):
    step = (K // J + 63) // 64 * 64
    acc = np.zeros((M, N))
    for k0 in range(0, K, step):
        acc += eng.debug_fc(A[:, k0:k0 + step], W[:, k0:k0 + step], b, 0, 1).astype(np.float64)
    stats('tc K split in %d (f64 sum)' % J, acc)
# emulate operand split error only (exact accumulation)
import torch
def split(x):
    t = torch.from_numpy(x); h = t.to(torch.bfloat16).float(); l = (t - h).to(torch.bfloat16).float(); return h.double().numpy(), l.double().numpy()
Ah, Al = split(A); Wh, Wl = split(W)
stats('emulated bf16x3 exact acc', Ah @ Wh.T + Ah @ Wl.T + Al @ Wh.T)
stats('emulated hi.hi only', Ah @ Wh.T)
