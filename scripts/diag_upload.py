"""Host->device volume upload of the bench workload: pageable (plain cudaMemcpyAsync vs the threaded pinned staging) and
pinned host arrays; then a cProfile of the end-to-end query (where the host time between the kernels goes)."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench as Bn, nnal_b200
eng = nnal_b200.get_engine()
model = Bn.make_model()
eng.set_model(model)
eng.volume_cache = False
for pinned, plain in ((False, 1), (False, 0), (True, 0)):
    padded, stats, pool = Bn.make_workload(100000, pinned=pinned)
    eng.debug_option('plain_upload', plain)
    for r in range(4):
        eng.synchronize(); t0 = time.perf_counter()
        eng.upload(0, padded)
        eng.synchronize(); t1 = time.perf_counter()
    nb = sum(a.nbytes for a in padded)
    print('pinned' if pinned else 'pageable', 'plain' if plain else 'staged', 'upload %.2f ms  %.1f GB/s' % (1e3 * (t1 - t0), nb / (t1 - t0) / 1e9))

class Expr(object):
    pass
expr = Expr()
expr.pars = dict(k=Bn.K_QUERY, B=10000, lambda_=0., patch_shape=Bn.PATCH, ntb=10000, stats=stats, fi_layers=2, fi_diag_load=Bn.FI_DELTA)
q = lambda: nnal_b200.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy+fi')
for _ in range(3):
    q()
eng.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    q()
eng.synchronize(); print('e2e query %.2f ms' % (1e2 * (time.perf_counter() - t0)))
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    q()
eng.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(22)
