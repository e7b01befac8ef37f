import sys, numpy as np
sys.path.insert(0, '.')
import nnal_b200, oracle as O
from tests.test_gpu_parity import _pw_setup, Expr
eng = nnal_b200.get_engine()
for (npool, seed) in [(700, 30), (384, 20), (1500, 31)]:
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(npool, seed)
    model = nnal_b200.NN.create_PW1(2); model.set_weights(w)
    expr = Expr(k=50, B=200, lambda_=0., patch_shape=ps, ntb=256, stats=stats)
    qo, posts = O.query_entropy_single(layers, w, padded, pool, ps, 256, stats, 50)
    for rep in range(4):
        for tc in (1, 0):
            eng.set_tensor_cores(tc)
            q = nnal_b200.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
            got = eng.pool_posteriors()[1]
            err = np.abs(got - posts)
            bad = np.where(err > 1e-4)[0]
            print('pool', npool, 'rep', rep, 'tc', tc, 'max err %.3g' % err.max(), 'n bad', len(bad), bad[:10], 'nan', np.isnan(got).sum())
    eng.set_tensor_cores(1)
