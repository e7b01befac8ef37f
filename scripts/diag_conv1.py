"""Timing experiments on the conv1 that gathers its own input (debug option wt_flags: 1 no normalise/split, 2 no operand
assembly, 4 no fetch): pool pass and per-layer times."""
import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import bench as Bn, nnal_b200
from nnal_b200 import _lib as L
padded, stats, pool = Bn.make_workload(100000)
model = Bn.make_model()
eng = nnal_b200.get_engine(); eng.set_model(model); eng.upload(0, padded)
st = np.array(stats, dtype=np.float64)
d_inds = torch.from_numpy(pool).cuda()
for fl in [int(a) for a in sys.argv[1:]] or [0]:
    eng.debug_option('wt_flags', fl)
    for r in range(3):
        eng.pool_begin(100000, 0)
        eng.profile(r == 2)
        eng.synchronize(); t0 = time.perf_counter()
        eng.pool_eval_device(0, d_inds.data_ptr(), 100000, 0, Bn.PATCH, st)
        eng.synchronize(); t1 = time.perf_counter()
    lay = [round(eng.profile_read(i)[0], 2) for i in range(9)]
    eng.profile(False)
    print('flags', fl, 'pool pass %.2f ms' % (1e3 * (t1 - t0)), 'layers', lay)
