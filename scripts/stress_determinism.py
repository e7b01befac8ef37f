"""Run every tensor-core kernel repeatedly on the same inputs; any run-to-run difference is a race."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nnal_b200
import oracle as O
eng = nnal_b200.get_engine()
R = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rs = np.random.RandomState(0)
for (H, Cin, Cout, ks) in [(25, 3, 24, 5), (25, 24, 32, 5), (13, 32, 48, 3), (13, 48, 96, 3)]:
    for n in (300, 1000):
        x = np.maximum(rs.randn(n, H, H, Cin), 0).astype(np.float32)
        W = (rs.randn(ks, ks, Cin, Cout) * np.sqrt(2. / (ks * ks * Cin))).astype(np.float32)
        b = (rs.randn(Cout) * .1).astype(np.float32)
        ref = eng.debug_conv(x, W, b, 1)
        bad = 0
        for r in range(R):
            o = eng.debug_conv(x, W, b, 1)
            if not np.array_equal(o, ref):
                bad += 1
                d = np.abs(o - ref)
                idx = np.argwhere(d > 0)
                print('   conv mismatch run', r, 'count', len(idx), 'max', d.max(), 'first', idx[:3].tolist(), 'samples', np.unique(idx[:, 0])[:10])
        print('conv H=%d Cin=%d Cout=%d n=%d: %d/%d runs differ' % (H, Cin, Cout, n, bad, R))
for (M, N, K) in [(1000, 4096, 4704), (777, 4096, 4096)]:
    A = np.maximum(rs.randn(M, K), 0).astype(np.float32)
    W = (rs.randn(N, K) * np.sqrt(2. / K)).astype(np.float32)
    b = (rs.randn(N) * .1).astype(np.float32)
    ref = eng.debug_fc(A, W, b, 1, 1)
    bad = sum(0 if np.array_equal(eng.debug_fc(A, W, b, 1, 1), ref) else 1 for _ in range(R // 3))
    print('fc M=%d N=%d K=%d: %d/%d runs differ' % (M, N, K, bad, R // 3))
# whole forward
from tests.util import pad_imgs, synth_volume, vol_stats
ps = (25, 25, 1)
imgs = synth_volume((40, 36, 6), 3, 30)
padded = pad_imgs(imgs, ps); stats = vol_stats(imgs)
pool = np.random.RandomState(31).choice(40 * 36 * 6, 700, replace=False).astype(np.int64)
layers = O.pw1_layers(2)
w = O.he_init_weights(layers, (25, 25, 3), 32)
model = nnal_b200.NN.create_PW1(2); model.set_weights(w)
ref = None; bad = 0
for r in range(R):
    posts = nnal_b200.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, ['posteriors'])[0]
    if ref is None: ref = posts
    elif not np.array_equal(posts, ref):
        bad += 1
        d = np.abs(posts - ref); print('   forward mismatch run', r, 'n', int((d > 0).sum()), 'max', d.max(), np.where(d > 0)[0][:10])
print('forward: %d/%d runs differ' % (bad, R))
