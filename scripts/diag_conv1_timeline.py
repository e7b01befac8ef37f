"""clock64 timeline of CTA 0 of the fused conv1 (debug option wt_flags & 256), one pool pass per flag set."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench as Bn, nnal_b200
padded, stats, pool = Bn.make_workload(20000)
eng = nnal_b200.get_engine(); eng.set_model(Bn.make_model()); eng.upload(0, padded)
st = np.array(stats, dtype=np.float64)
d_inds = torch.from_numpy(pool).cuda()
fl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
eng.debug_option('wt_flags', 256 | fl)
eng.pool_begin(20000, 0)
eng.pool_eval_device(0, d_inds.data_ptr(), 20000, 0, Bn.PATCH, st)
eng.synchronize()
