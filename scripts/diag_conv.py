import sys, numpy as np
sys.path.insert(0, '.')
import nnal_b200, oracle as O
eng = nnal_b200.get_engine()
for (H, Cin, Cout, ks) in [(25, 24, 32, 5), (13, 32, 48, 3), (13, 48, 96, 3)]:
    for n in [149, 300, 2000]:
        rs = np.random.RandomState(1)
        x = np.maximum(rs.randn(n, H, H, Cin), 0).astype(np.float32)
        W = (rs.randn(ks, ks, Cin, Cout) * np.sqrt(2. / (ks * ks * Cin))).astype(np.float32)
        b = (rs.randn(Cout) * .1).astype(np.float32)
        ref = eng.debug_conv(x, W, b, 0)
        got = eng.debug_conv(x, W, b, 1)
        err = np.abs(got - ref).reshape(n, -1).max(1) / np.abs(ref).max()
        bad = np.where(err > 1e-4)[0]
        print('shape', (H, Cin, Cout, ks), 'n', n, 'bad samples', len(bad), bad[:20], 'max err', err.max())
        if len(bad):
            s = bad[0]
            e = np.abs(got[s] - ref[s]).max(-1) > 1e-4 * np.abs(ref).max()
            print(' first bad sample', s, 'bad positions', int(e.sum()), 'rows', np.where(e.any(1))[0][:30], 'cols', np.where(e.any(0))[0][:30])
            ch = np.abs(got[s] - ref[s]).reshape(-1, Cout).max(0) > 1e-4 * np.abs(ref).max()
            print(' bad channels', np.where(ch)[0])
