"""Where the host-side jitter of the end-to-end query comes from: per-call wall times (with a synchronise after each) of the
engine calls one 'entropy+fi' query makes, for 40 queries right after process start."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as Bn, nnal_b200
eng = nnal_b200.get_engine()
model = Bn.make_model()
padded, stats, pool = Bn.make_workload(100000, pinned=True)
eng.volume_cache = False
names = ['set_model', 'set_volume', 'pool_begin', 'pool_eval', 'pool_score', 'pool_topk', 'fi_set_candidates', 'fi_begin', 'fi_greedy', 'fi_info']
log = {}
sync = '--nosync' not in sys.argv
def wrap(name):
    f = getattr(eng, name)
    def g(*a, **k):
        t0 = time.perf_counter()
        r = f(*a, **k)
        if sync:
            eng.synchronize()
        log[name] = log.get(name, 0.) + 1e3 * (time.perf_counter() - t0)
        return r
    setattr(eng, name, g)
for n_ in names:
    if hasattr(eng, n_):
        wrap(n_)
class Expr(object):
    pass
expr = Expr()
expr.pars = dict(k=Bn.K_QUERY, B=10000, lambda_=0., patch_shape=Bn.PATCH, ntb=10000, stats=stats, fi_layers=2, fi_diag_load=Bn.FI_DELTA)
for it in range(40):
    log.clear()
    t0 = time.perf_counter()
    nnal_b200.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy+fi')
    tot = 1e3 * (time.perf_counter() - t0)
    print('%2d total %6.1f  other %5.1f  ' % (it, tot, tot - sum(log.values())) + ' '.join('%s %.1f' % (k, v) for k, v in log.items() if v > 0.3))
