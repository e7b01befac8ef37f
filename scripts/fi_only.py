"""Small FI round used for the ncu launch list and timing experiments: random factors (n candidates, d = d_prev = 4096),
greedy k, Gram.  Optional third argument: nnal_debug_option('fi_flags', v) (8 step of round 1: inverse inside the column kernel, 16 pipelined step without speculative columns; 1/2/4: timing experiments)."""
import sys, os, time, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nnal_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
k = int(sys.argv[2]) if len(sys.argv) > 2 else 100
flags = [int(a) for a in sys.argv[3:]] or [0]
rs = np.random.RandomState(0)
U = np.maximum(rs.randn(n, 4096), 0).astype(np.float32)
A = np.maximum(rs.randn(n, 4096), 0).astype(np.float32)
Wl = (rs.randn(2, 4096) * .02).astype(np.float32)
p1 = rs.rand(n)
eng = nnal_b200.get_engine()
eng.fi_set_factors(p1, U, A, Wl)
for fl in flags:
    eng.debug_option('fi_flags', fl)
    for it in range(3):
        eng.synchronize(); t0 = time.perf_counter()
        sel, obj, red = eng.fi_greedy(k, 1e-5)
        eng.synchronize(); t1 = time.perf_counter()
        print('flags %d: greedy %.2f ms (%.1f us/step)  sel %s  red[-1] %.12g' % (fl, 1e3 * (t1 - t0), 1e6 * (t1 - t0) / k,
              hashlib.sha1(np.asarray(sel, dtype=np.int64).tobytes()).hexdigest()[:12], red[-1]))
eng.debug_option('fi_flags', 0)
eng.synchronize(); t1 = time.perf_counter()
eng.fi_gram(None, read=False)
eng.synchronize(); t2 = time.perf_counter()
print('gram %.2f ms' % (1e3 * (t2 - t1)))
