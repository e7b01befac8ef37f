"""Pool-pass time of the bench workload against the forward chunk size (debug option `chunk`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as Bn, nnal_b200
padded, stats, pool = Bn.make_workload(100000)
eng = nnal_b200.get_engine(); eng.set_model(Bn.make_model()); eng.upload(0, padded)
st = np.array(stats, dtype=np.float64)
d_inds = torch.from_numpy(pool).cuda()
for ch in [int(a) for a in sys.argv[1:]] or [16384]:
    eng.debug_option('chunk', ch)
    ts = []
    for r in range(6):
        eng.pool_begin(100000, 2)
        eng.synchronize(); t0 = time.perf_counter()
        eng.pool_eval_device(0, d_inds.data_ptr(), 100000, 0, Bn.PATCH, st)
        eng.synchronize(); ts.append(1e3 * (time.perf_counter() - t0))
    print('chunk %6d: pool pass %.2f ms (min of %s)' % (ch, min(ts[1:]), ' '.join('%.2f' % t for t in ts[1:])))
