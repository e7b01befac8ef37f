"""Parity at BASELINE scale (``pytest -m gpu``): config 2's full 100,000-patch pool on full-size volumes and config 1's full
2,000-sample pool, against the float64 restatement evaluated with torch on the host cores (oracle/torch_fp32.py,
dtype=float64: same semantics as oracle.forward, pinned against it below on a sample)."""
import os

import numpy as np
import pytest

import oracle as O
from tests.util import assert_topk_equivalent

pytestmark = pytest.mark.gpu


class Expr(object):
    def __init__(self, **pars):
        self.pars = pars


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


def _torch64(layers, w, fl):
    import torch
    from oracle.torch_fp32 import TorchForward
    return TorchForward(layers, w, feature_layer=fl, threads=os.cpu_count(), dtype=torch.float64)


def test_config2_full_pool_100k(nb):
    """BASELINE config 2 as bench.py runs it: 3 x 256x256x180 volumes, 100,000-patch pool, PW1 c = 2, He-normal weights.
    Posteriors of EVERY pool sample within 1e-4 of float64; the 'entropy' answer (k = 100) is a valid top-100 of the
    float64 scores up to ties inside that tolerance; the FI pre-filter's B = 10,000 likewise."""
    import bench as Bn
    padded, stats, pool = Bn.make_workload(100000)
    layers = O.pw1_layers(2)
    w = O.he_init_weights(layers, (25, 25, 3), 4)
    model = Bn.make_model()
    assert all(np.array_equal(model.var_dict[k][0], w[k][0]) for k in w)          # bench weights == oracle weights
    fwd = _torch64(layers, w, len(layers) - 2)
    # the fast float64 forward is the oracle's forward (sample check)
    xs = O.normalize_batch_eval(O.get_patches(padded, pool[:24], Bn.PATCH), stats).astype(np.float32)
    assert np.abs(fwd(xs)['posteriors'] - O.forward(layers, w, xs)['posteriors']).max() < 1e-12
    posts = O.batch_eval(layers, w, padded, pool, Bn.PATCH, 2000, stats, 'posteriors', fwd=fwd)[0]
    expr = Expr(k=100, B=10000, lambda_=0., patch_shape=Bn.PATCH, ntb=10000, stats=stats, fi_layers=2, fi_diag_load=1e-5)
    q_ent, q_fi = nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy+fi')
    got = nb.get_engine().pool_posteriors()[1].astype(np.float64)
    err = np.abs(got - posts).max()
    assert err < 1e-4, 'max posterior error %g over 100k patches' % err
    score = np.abs(posts - .5)
    assert_topk_equivalent(q_ent, score, 100, 1e-4)
    # the FI selection lives inside the 10,000 most uncertain samples (up to ties at the boundary within tolerance)
    kth = np.sort(score)[9999]
    assert np.all(score[q_fi] <= kth + 1e-4) and len(np.unique(q_fi)) == 100
    q_only = nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
    assert np.array_equal(q_only, q_ent)


def test_config1_full_pool_2000(nb):
    """BASELINE config 1: the reference small CNN (PW1 layer dict on 28x28x1, c = 10) over a 2,000-sample pool: posteriors
    within 1e-4, entropy query k = 10 (NNAL.py:298-310) and the last-layer FI trace ranking (NNAL.py:121-139) valid up to
    ties inside the tolerance."""
    from collections import OrderedDict
    layers = O.pw1_layers(10)
    w = O.he_init_weights(layers, (28, 28, 1), 1)
    x = np.random.RandomState(0).rand(2000, 28, 28, 1).astype(np.float32)
    model = nb.NN.CNN((28, 28, 1), OrderedDict(layers), feature_layer=len(layers) - 2)
    model.set_weights(w)
    r = _torch64(layers, w, len(layers) - 2)(x)
    post, feat = r['posteriors'], r['feature_layer']
    expr = Expr(k=10, B=100, lambda_=0., batch_size=500)
    expr.pool_images = x
    q = nb.NNAL.CNN_query(model, expr, np.arange(2000), 'entropy', None)
    got = nb.get_engine().pool_posteriors().astype(np.float64)
    assert np.abs(got - post).max() < 1e-4
    # the 28 x 28 / 14 x 14 conv shapes of this model have their own tcgen05 tile plans (conv_tc.cu CfgS1Conv1-4): no layer
    # with weights falls back to the FP32 CUDA-core kernels
    eng = nb.get_engine()
    for i, (name, spec) in enumerate(layers[:-1]):
        if spec[1] in ('conv', 'fc'):
            assert eng.layer_info(i)[2] != 0, '%s runs on CUDA cores' % name
    H = O.compute_entropy(post.copy())
    assert_topk_equivalent(q, -H, 10, 1e-3 * np.abs(H).max())
    q = nb.NNAL.CNN_query(model, expr, np.arange(2000), 'fi', None)
    tr = O.fi_trace_score(post, feat)
    assert_topk_equivalent(q, -tr, 10, 1e-3 * np.abs(tr).max())
