"""Bring-up helper for conv_wt.cu: error maps of the weight-stationary conv vs the float64 oracle."""
import sys
import numpy as np
sys.path.insert(0, '.')
import nnal_b200
import oracle as O
eng = nnal_b200.get_engine()
for (H, Cin, Cout, ks, mode) in [(25, 24, 32, 5, 2), (25, 24, 32, 5, 3), (25, 3, 24, 5, 2), (13, 32, 48, 3, 2)]:
    for n in (1, 3, 300):
        rs = np.random.RandomState(n)
        x = np.maximum(rs.randn(n, H, H, Cin), 0).astype(np.float32)
        W = (rs.randn(ks, ks, Cin, Cout) * np.sqrt(2. / (ks * ks * Cin))).astype(np.float32)
        b = (rs.randn(Cout) * .1).astype(np.float32)
        ref = np.maximum(O.conv2d_same(x.astype(np.float64), W.astype(np.float64), b.astype(np.float64)), 0)
        if mode == 3:
            ref = O.max_pool_same(ref)
        try:
            got = eng.debug_conv(x, W, b, mode)
        except Exception as e:
            print(H, Cin, Cout, 'mode', mode, 'n', n, 'FAILED', e)
            break
        e = np.abs(got - ref) / np.abs(ref).max()
        print(H, Cin, Cout, 'mode', mode, 'n', n, 'max rel err %.3g' % e.max())
        if e.max() > 1e-4:
            bad = e > 1e-4
            print('  bad fraction %.3f; by sample' % bad.mean(), bad.mean(axis=(1, 2, 3))[:6])
            print('  by y', np.round(bad.mean(axis=(0, 2, 3)), 2))
            print('  by x', np.round(bad.mean(axis=(0, 1, 3)), 2))
            print('  by c', np.round(bad.mean(axis=(0, 1, 2)), 2))
