"""Fine-grained host timing of the FI round stages (run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as Bn
import nnal_b200
from nnal_b200 import _lib as L

padded, stats, pool = Bn.make_workload(100000)
layers, w = Bn.pw1_weights()
model = nnal_b200.NN.create_PW1(2); model.set_weights(w)
eng = nnal_b200.get_engine(); eng.set_model(model); eng.upload(0, padded)
st = np.array(stats)
B, k = 10000, 100

def T(label, f):
    eng.synchronize(); t0 = time.perf_counter(); r = f(); eng.synchronize()
    print('%-28s %8.3f ms' % (label, 1e3 * (time.perf_counter() - t0))); return r

for it in range(3):
    print('--- iteration', it)
    T('pool_begin(n,0)', lambda: eng.pool_begin(len(pool), 0))
    T('pool_eval', lambda: eng.pool_eval(0, pool, 0, Bn.PATCH, st, shape=padded[0].shape))
    T('score', lambda: eng.pool_score(L.SCORE_BINARY))
    idx, sc = T('topk B', lambda: eng.pool_topk(B, with_scores=True))
    T('pool_begin(B,2)', lambda: eng.pool_begin(B, 2))
    T('pool_eval B', lambda: eng.pool_eval(0, pool[idx], 0, Bn.PATCH, st, shape=padded[0].shape))
    T('fi_set_candidates', lambda: eng.fi_set_candidates(None, 2))
    T('fi_greedy', lambda: eng.fi_greedy(k, 1e-5))
    T('fi_gram', lambda: eng.fi_gram(None, read=False))
