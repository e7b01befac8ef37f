"""Parity check of the literal FI pipeline against the float64 oracle on a small PW1 pool.  tests/test_gpu_sdp.py calls
``check()`` with the test-only kernel-selection switches (``Engine.debug_option``: csrc/shrunk.cu / csrc/sdp.cu fallbacks)
set, so that the fallback kernels (fp32 forward, fp32 fc gradient, filter through L2, 4-channel register tile, one launch
per SDP iteration) are exercised on the same inputs.  Stand-alone: prints 'OK <objective>' or raises."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def check():
    import nnal_b200
    import oracle as O
    from tests.test_gpu_fi import _pw_setup
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(40, 90)
    model = nnal_b200.NN.create_PW1(2)
    model.set_weights(w)
    eng = nnal_b200.get_engine()
    eng.set_model(model, None)
    eng.upload(0, padded)
    post, g = eng.fi_shrunk_voxels(0, pool, ps, np.array(stats, dtype=np.float64), shape=padded[0].shape)
    x = O.normalize_batch_eval(O.get_patches(padded, pool, ps), stats).astype(np.float32)
    po, go = O.shrunk_class_gradients(layers, w, x)
    floor = 1e-3 * np.abs(go).max()
    for t in range(go.shape[2]):
        scale = max(np.abs(go[:, :, t]).max(), floor)
        err = np.abs(g[:, :, t] - go[:, :, t]).max()
        assert err <= 2e-4 * scale, 'layer %d: %g vs %g' % (t, err, scale)
    A = np.array(O.gen_A_matrices(go[0], go[1], po[1], 1e-5))
    r = eng.sdp_query_distribution(A, tol=1e-4)
    phi, gap = O.sdp_certificate(A, r['q'])
    assert abs(r['objective'] / phi - 1) < 1e-9 and gap <= 2e-4, (r['objective'], phi, gap)
    qo, to, phio, gapo, ito = O.sdp_solve(A, 1e-4)
    assert abs(phi / phio - 1) < 1e-3
    return r


if __name__ == '__main__':
    r = check()
    print('OK %.12e %d' % (r['objective'], r['iterations']))
