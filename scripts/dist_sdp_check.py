"""Multi-GPU check of the literal FI pipeline (fi_mode='sdp'): every rank back-propagates its block of the B candidates,
the shrunk gradients are all-gathered, the SDP is solved on every rank, rank 0's draw is broadcast.
  python scripts/dist_sdp_check.py                       (one process)
  torchrun --nproc-per-node 2 scripts/dist_sdp_check.py  (one process per GPU)"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
if world > 1:
    import torch
    import torch.distributed as td
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    td.init_process_group('nccl')
import nnal_b200
import oracle as O


class Expr(object):
    pass


rs = np.random.RandomState(0)
shape = (96, 96, 8)
imgs = [np.clip(rs.randn(*shape) * 30 + 100, 0, None).astype(np.float32) for _ in range(3)]
padded = [np.pad(im, ((12, 12), (12, 12), (0, 0)), 'constant') for im in imgs]
stats = [[float(im.mean()), float(im.std())] for im in imgs]
pool = rs.choice(int(np.prod(shape)), 20000, replace=False).astype(np.int64)
model = nnal_b200.NN.create_PW1(2)
model.set_weights(O.he_init_weights(O.pw1_layers(2), (25, 25, 3), 4, bias_scale=0.05))
expr = Expr()
expr.pars = dict(k=50, B=4096, lambda_=0., patch_shape=(25, 25, 1), ntb=10000, stats=stats, fi_mode='sdp')
expr.nclass = 2
for it in range(2):
    np.random.seed(11 + 100 * rank)          # different generator states: the broadcast must reconcile the draws
    t0 = time.perf_counter()
    q, soln, sel = nnal_b200.fi.query_single_sdp(expr, model, None, padded, pool, return_solution=True)
    nnal_b200.get_engine().synchronize()
    dt = time.perf_counter() - t0
print('rank %d/%d: %.1f ms, objective %.9e, gap %.2e, iterations %d, q[:8] %s, checksum %d'
      % (rank, world, 1e3 * dt, soln['primal objective'], soln['gap'], soln['iterations'], q[:8].tolist(), int(q.sum())))
if world > 1:
    td.barrier()
    td.destroy_process_group()
