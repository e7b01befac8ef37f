#!/usr/bin/env python
"""ONE kind of work only, for ncu: R query rounds (entropy + FI, as bench.py's timed step) on resident inputs.

  python scripts/round_only.py [--pool 100000] [--fi-B 10000] [--rounds 2]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as Bn          # noqa: E402
import nnal_b200            # noqa: E402
from nnal_b200 import _lib as L, dist, fi as fimod     # noqa: E402
import torch                # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--pool', type=int, default=100000)
ap.add_argument('--fi-B', type=int, default=10000)
ap.add_argument('--rounds', type=int, default=2)
args = ap.parse_args()
padded, stats, pool = Bn.make_workload(args.pool)
model = Bn.make_model()
eng = nnal_b200.get_engine()
eng.set_model(model)
eng.upload(0, padded)
st = np.array(stats, dtype=np.float64)
d_inds = torch.from_numpy(pool).cuda()
n, k, B = args.pool, Bn.K_QUERY, min(args.fi_B, args.pool)
for r in range(args.rounds):
    eng.pool_begin(n, 2)
    eng.pool_eval_device(0, d_inds.data_ptr(), n, 0, Bn.PATCH, st)
    eng.pool_score(L.SCORE_BINARY)
    top, _ = dist.topk_global(eng, max(B, k), 0, n)
    eng.fi_set_candidates(top[:B], 2)
    eng.synchronize()
    t0 = time.perf_counter()
    chosen, obj, red = fimod.greedy_select(eng, k, Bn.FI_DELTA, np.arange(B, dtype=np.int64))
    eng.synchronize()
    print('greedy %.2f ms' % (1e3 * (time.perf_counter() - t0)))
eng.synchronize()
print('rounds', args.rounds, 'launches', eng.launches, 'entropy', Bn.digest(top[:k]), 'fi', Bn.digest(top[:B][chosen]), 'red', red[-1])
