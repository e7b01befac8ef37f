"""The reference's literal FI selection on B synthetic PW1 candidates, used for timing and the ncu launch list:
shrunk class-score gradients (csrc/shrunk.cu) -> A-matrices -> SDP query distribution (csrc/sdp.cu).
  python scripts/sdp_only.py [B] [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import nnal_b200
from nnal_b200.PW_NNAL import _A_from_shrunk

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
rs = np.random.RandomState(0)
shape = (96, 96, 8) if B <= 60000 else (128, 128, 16)
imgs = [np.clip(rs.randn(*shape) * 30 + 100, 0, None).astype(np.float32) for _ in range(3)]
padded = [np.pad(im, ((12, 12), (12, 12), (0, 0)), 'constant') for im in imgs]
stats = np.array([[im.mean(), im.std()] for im in imgs], dtype=np.float64)
pool = rs.choice(int(np.prod(shape)), B, replace=False).astype(np.int64)
model = nnal_b200.NN.create_PW1(2)
import oracle as O
model.set_weights(O.he_init_weights(O.pw1_layers(2), (25, 25, 3), 4, bias_scale=0.05))
eng = nnal_b200.get_engine()
eng.set_model(model, None)
eng.upload(0, padded)
for it in range(reps):
    eng.profile(True)
    eng.synchronize(); t0 = time.perf_counter()
    post, g = eng.fi_shrunk_voxels(0, pool, (25, 25, 1), stats, shape=padded[0].shape)
    t1 = time.perf_counter()
    fwd, _ = eng.profile_read(120)
    bwd, _ = eng.profile_read(121)
    eng.profile(False)
    A = _A_from_shrunk(g, post[1].astype(np.float64), 1e-5, as_list=False)
    t2 = time.perf_counter()
    r = eng.sdp_query_distribution(A, tol=1e-4)
    t3 = time.perf_counter()
    print('B %d: shrunk gradients %.2f ms (device: forward %.2f, backward %.2f), A %.2f ms, SDP %.2f ms (%d iterations, '
          '%.1f us each, gap %.2e, support %d)' % (B, 1e3 * (t1 - t0), fwd, bwd, 1e3 * (t2 - t1), 1e3 * (t3 - t2),
                                                   r['iterations'], 1e6 * (t3 - t2) / max(1, r['iterations']), r['gap'],
                                                   int((r['q'] > 1e-8).sum())))
