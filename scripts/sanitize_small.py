"""Small run of the kernels changed late in round 2, for compute-sanitizer (memcheck / racecheck): the fused conv1 pool pass
(ragged last chunk) and the tensor-core conv data gradients."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import nnal_b200
import oracle as O
from tests.util import centered_weights, pad_imgs, synth_volume, vol_stats
ps = (25, 25, 1)
imgs = synth_volume((40, 36, 5), 3, 0)
padded = pad_imgs(imgs, ps)
stats = vol_stats(imgs)
pool = np.random.RandomState(1).choice(40 * 36 * 5, 700, replace=False).astype(np.int64)
layers = O.pw1_layers(2)
probe = O.normalize_batch_eval(O.get_patches(padded, pool[:32], ps), stats).astype(np.float32)
w = centered_weights(layers, (25, 25, 3), 2, probe)
model = nnal_b200.NN.create_PW1(2)
model.set_weights(w)
eng = nnal_b200.get_engine()
eng.debug_option('chunk', 300)
eng.debug_option('bw_chunk', 100)
p = nnal_b200.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
want = O.batch_eval(layers, w, padded, pool, ps, 100, stats, 'posteriors')[0]
print('posteriors max err', np.abs(p - want).max())
eng.set_model(model, None)
eng.upload(0, padded)
post, g = eng.fi_shrunk_voxels(0, pool[:250], ps, np.array(stats, dtype=np.float64), shape=padded[0].shape)
print('shrunk ok', g.shape, float(np.abs(g).max()))
