"""GPU parity of the last-layer influence pieces: NN.LLFC_grads / NN.LLFC_hess and
PW_NNAL.stoch_approx_IF against the oracle, which is pinned to the unmodified reference (tests/golden: if_V)."""
from collections import OrderedDict

import numpy as np
import pytest

import oracle as O
from tests.test_reference_dispatch_golden import LAYERS_W

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


def _model(nb):
    w = O.he_init_weights(LAYERS_W, (9, 7, 2), 6, bias_scale=0.1)
    model = nb.NN.CNN((9, 7, 2), OrderedDict(LAYERS_W), feature_layer=len(LAYERS_W) - 2)
    model.set_weights(w)
    return model, w


def test_stoch_approx_IF_matches_reference_golden(nb, golden):
    """Same pool / training patches / np.random seed as the run of the UNMODIFIED reference that produced the golden V."""
    model, w = _model(nb)
    x = golden['w_pool']
    tr_x, pl_x = x[:12], x[40:47]
    np.random.seed(31)
    V, lab = nb.PW_NNAL.stoch_approx_IF(model, None, tr_x, pl_x, 25, 20)
    assert np.array_equal(lab, golden['if_labels'])
    ref = golden['if_V']
    assert V.shape == ref.shape == ((24 + 1) * 3, 7)
    # float32 device forward vs the float64 restatement behind the golden: features agree to ~1e-6
    assert np.abs(V - ref).max() < 2e-5 * np.abs(ref).max()
    # and exactly the recursion of the oracle on the DEVICE's own factors (float64 arithmetic both sides)
    post, U = nb.NN._last_layer_factors(model, None, {model.x: pl_x})
    tp, tU = nb.NN._last_layer_factors(model, None, {model.x: tr_x})
    Vo, _ = O.stoch_approx_IF(post.astype(np.float64), U.astype(np.float64), tp.astype(np.float64), tU.astype(np.float64),
                              list(golden['if_draws']), 20.)
    assert np.abs(V - Vo).max() < 1e-6 * np.abs(Vo).max()      # (the reference rounds pi*u to float32; kept on the device)


@pytest.mark.parametrize('n,c,d,T', [(5, 2, 64, 40), (33, 3, 100, 7), (3, 10, 4096, 12), (4, 2, 4096, 0)])
def test_if_lissa_factored_equals_dense(nb, n, c, d, T):
    """nnal_if_lissa (factored, V slice in shared or global memory) == the dense recursion with explicit Hessians."""
    rs = np.random.RandomState(n + c + d)
    z = rs.randn(c, n)
    post = (np.exp(z) / np.exp(z).sum(0)).astype(np.float32)
    U = np.maximum(rs.randn(n, d), 0).astype(np.float32)
    lab = rs.randint(0, c, n)
    zt = rs.randn(max(T, 1), c)
    tp = (np.exp(zt) / np.exp(zt).sum(1, keepdims=True)).astype(np.float32)[:T]
    tU = np.maximum(rs.randn(max(T, 1), d), 0).astype(np.float32)[:T] * 0.1
    V = nb.get_engine().if_lissa(post, U, lab, tp, tU, 30.)
    G = O.LLFC_grads(post, U.T, lab)                          # float32 inputs: the same float32 pi*u product as upstream
    Vo = G.astype(np.float64)
    if d <= 128:
        for t in range(T):
            H = -O.LLFC_hess(tp[t:t + 1].T.astype(np.float64), tU[t:t + 1].T.astype(np.float64))
            Vo = G + Vo - H @ Vo / 30.
    else:                                                     # factored float64 NumPy (the dense Hessian would be 13 GB)
        for t in range(T):
            ut = np.append(tU[t].astype(np.float64), 1.)
            pt = tp[t].astype(np.float64)
            P = np.diag(pt) - np.outer(pt, pt)
            Vm = np.concatenate([Vo[:c * d].reshape(c, d, n), Vo[c * d:].reshape(c, 1, n)], axis=1)      # [c][d+1][n]
            s = np.einsum('akn,k->an', Vm, ut)
            r = P @ s
            HV = np.einsum('an,k->akn', r, ut)
            HVf = np.concatenate([HV[:, :d].reshape(c * d, n), HV[:, d]], axis=0)
            Vo = G + Vo - HVf / 30.
    assert V.shape == ((d + 1) * c, n)
    assert np.abs(V - Vo).max() <= 1e-9 * max(np.abs(Vo).max(), 1e-300)


def test_llfc_grads_and_hess(nb, golden):
    model, w = _model(nb)
    x = golden['w_pool'][:9]
    r = O.forward(LAYERS_W, w, x, feature_layer=len(LAYERS_W) - 2)
    G, lab = nb.NN.LLFC_grads(model, None, {model.x: x})
    Go, labo = O.LLFC_grads(r['posteriors'], r['feature_layer'])
    assert np.array_equal(lab, labo) and np.abs(G - Go).max() < 1e-5 * np.abs(Go).max()
    lab2 = np.arange(9) % 3
    assert np.abs(nb.NN.LLFC_grads(model, None, {model.x: x}, lab2) - O.LLFC_grads(r['posteriors'], r['feature_layer'], lab2)).max() < 1e-5
    H = nb.NN.LLFC_hess(model, None, {model.x: x[4:5]})
    Ho = O.LLFC_hess(r['posteriors'][:, 4:5], r['feature_layer'][:, 4:5])
    assert H.shape == Ho.shape and np.abs(H - Ho).max() < 1e-5 * np.abs(Ho).max()
