"""Small FI round used for the ncu launch list: random factors (n candidates, d = d_prev = 4096), greedy k, Gram."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nnal_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rs = np.random.RandomState(0)
U = np.maximum(rs.randn(n, 4096), 0).astype(np.float32)
A = np.maximum(rs.randn(n, 4096), 0).astype(np.float32)
Wl = (rs.randn(2, 4096) * .02).astype(np.float32)
p1 = rs.rand(n)
eng = nnal_b200.get_engine()
eng.fi_set_factors(p1, U, A, Wl)
for it in range(2):
    eng.synchronize(); t0 = time.perf_counter()
    sel, obj, red = eng.fi_greedy(k, 1e-5)
    eng.synchronize(); t1 = time.perf_counter()
    eng.fi_gram(None, read=False)
    eng.synchronize(); t2 = time.perf_counter()
    print('greedy %.2f ms (%.1f us/step), gram %.2f ms' % (1e3 * (t1 - t0), 1e6 * (t1 - t0) / k, 1e3 * (t2 - t1)))
