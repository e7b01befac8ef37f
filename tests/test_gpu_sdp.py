"""GPU parity tests of the reference's literal FI pipeline (``pytest -m gpu`` on the B200 box): shrunk class-score
gradients (NN.get_gradients + NNAL_tools.shrink_gradient), gen_A_matrices (row 10), the SDP query
distribution (row 12) and the 'fi' query with ``fi_mode='sdp'`` -- through the C ABI, against oracle/fi_oracle.py."""
from collections import OrderedDict

import numpy as np
import pytest

import oracle as O
from tests.test_gpu_fi import Expr, _pw_setup
from tests.test_gpu_parity import SMALL

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-3      # north_star: FI objectives within 1e-3 relative
G_RTOL = 2e-4        # shrunk gradients: fp32 forward/backward vs the float64 oracle, relative to the layer's largest entry

SMALL2 = [('conv1', [8, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']), ('conv2', [12, 'conv', [5, 3]]),
          ('fc1', [72, 'fc']), ('fc2', [64, 'fc']), ('fc3', [2, 'fc'])]


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


def _assert_shrunk_close(g, go, rtol=G_RTOL, outliers=0.0):
    """``outliers``: fraction of samples allowed to miss ``rtol`` (by at most 100x): a ReLU unit or a max-pool window whose
    float32 activation sits within round-off of 0 / of a tie takes the other branch than in float64 -- any float32
    implementation, TF included, differs from the float64 oracle there."""
    assert g.shape == go.shape
    floor = 1e-3 * np.abs(go).max()      # the last layer's entry is identically 0 (sum_y dz_y = 0): round-off only
    for t in range(go.shape[2]):
        scale = max(np.abs(go[:, :, t]).max(), floor)
        err = np.abs(g[:, :, t] - go[:, :, t])
        if outliers > 0:
            bad = (err > rtol * scale).any(axis=0)
            assert bad.mean() <= outliers and err.max() <= 100 * rtol * scale, \
                'layer %d: %d samples off, max %g vs scale %g' % (t, bad.sum(), err.max(), scale)
            continue
        assert np.abs(g[:, :, t] - go[:, :, t]).max() <= rtol * scale, \
            'layer %d: %g vs scale %g' % (t, np.abs(g[:, :, t] - go[:, :, t]).max(), scale)


@pytest.mark.parametrize('layers,in_shape,seed', [(SMALL, (9, 7, 2), 3), (SMALL2, (11, 8, 3), 4)])
@pytest.mark.parametrize('tc', [1, 0])
def test_shrunk_gradients_small_nets(nb, layers, in_shape, seed, tc):
    """Generic layer dictionaries (odd sizes, non-square filters, c = 3 -> one backward pass per class; c = 2 -> the
    single-pass shortcut) == shrink_gradient(tf.gradients(log p_y), 'sum') of the float64 oracle."""
    rs = np.random.RandomState(seed)
    x = rs.randn(150, *in_shape).astype(np.float32)
    w = O.he_init_weights(layers, in_shape, seed, bias_scale=0.1)
    model = nb.NN.CNN(in_shape, OrderedDict(layers), feature_layer=len(layers) - 2)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_model(model, None)
    eng.set_tensor_cores(tc)
    try:
        post, g = eng.fi_shrunk_images(x)
    finally:
        eng.set_tensor_cores(1)
    po, go = O.shrunk_class_gradients(layers, w, x)
    assert eng.fi_shrunk_tau() == go.shape[2]
    assert np.abs(post - po).max() < 1e-4
    _assert_shrunk_close(g, go)


def test_shrunk_gradients_chunking(nb, monkeypatch):
    """Results do not depend on the chunk size of the backward pass."""
    rs = np.random.RandomState(8)
    x = rs.randn(70, 9, 7, 2).astype(np.float32)
    w = O.he_init_weights(SMALL, (9, 7, 2), 9, bias_scale=0.1)
    model = nb.NN.CNN((9, 7, 2), OrderedDict(SMALL), feature_layer=len(SMALL) - 2)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_model(model, None)
    p1, g1 = eng.fi_shrunk_images(x)
    eng.debug_option('bw_chunk', 16)
    try:
        p2, g2 = eng.fi_shrunk_images(x)
    finally:
        eng.debug_option('bw_chunk', 0)
    assert np.array_equal(g1, g2) and np.array_equal(p1, p2)
    p3, g3 = eng.fi_shrunk_images(x[:0])
    assert g3.shape == (3, 0, 6)


def test_shrunk_gradients_pw1_and_A_matrices(nb):
    """PW1 on gathered voxels: shrunk gradients, then PW_NNAL.gen_A_matrices == the oracle's A_i (rows 8-10)."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(48, 90)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_model(model, None)
    eng.upload(0, padded)
    st = np.array(stats, dtype=np.float64)
    post, g = eng.fi_shrunk_voxels(0, pool, ps, st, shape=padded[0].shape)
    x = O.normalize_batch_eval(O.get_patches(padded, pool, ps), stats).astype(np.float32)
    po, go = O.shrunk_class_gradients(layers, w, x)
    assert go.shape == (2, 48, 7)
    assert np.abs(post - po).max() < 1e-4
    _assert_shrunk_close(g, go)
    # host-patch entry point (the reference hands gen_A_matrices the normalised patches, PW_NNAL.py:121-137)
    sel_posts = po[1].copy()
    sel_posts[0], sel_posts[1] = 1e-9, 1 - 1e-9          # the two clamped branches (:770-793)
    expr = Expr(k=5, B=48, lambda_=0., patch_shape=ps, ntb=128, stats=stats)
    A = nb.PW_NNAL.gen_A_matrices(expr, model, None, x, sel_posts, 1e-5)
    Ao = O.gen_A_matrices(go[0], go[1], sel_posts, 1e-5)
    assert len(A) == 48
    for a, b in zip(A, Ao):
        assert np.abs(a - b).max() <= 5e-4 * np.abs(b - 1e-5 * np.eye(7)).max() + 1e-20
    # objective of the uniform design in both sets of matrices
    q = np.ones(48) / 48
    assert abs(O.sdp_objective(A, q) / O.sdp_objective(Ao, q) - 1) < OBJ_RTOL


def test_conv_data_gradient_tensor_core_vs_cuda_core(nb):
    """PW1's conv2 / conv3 / conv4 data gradients run as tcgen05 shift-GEMM convolutions of dz with the flipped, transposed
    filter (conv_tc.cu, fp16 hi/lo planes of the power-of-two scaled gradient); debug option bw_no_tc = 1 selects the fp32
    CUDA-core kernels (and the fp32 fc gradient).  Both meet the oracle's bar and agree with each other, ragged chunk included."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(150, 91)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_model(model, None)
    eng.upload(0, padded)
    st = np.array(stats, dtype=np.float64)
    eng.debug_option('bw_chunk', 64)
    try:
        l0 = eng.launches
        post, g = eng.fi_shrunk_voxels(0, pool, ps, st, shape=padded[0].shape)
        l_tc = eng.launches - l0
        eng.debug_option('bw_no_tc', 1)
        l0 = eng.launches
        post2, g2 = eng.fi_shrunk_voxels(0, pool, ps, st, shape=padded[0].shape)
        l_simt = eng.launches - l0
    finally:
        eng.debug_option('bw_no_tc', 0)
        eng.debug_option('bw_chunk', 0)
    assert l_tc > l_simt                    # (absmax + split launches of the three conv layers: the tensor-core path did run)
    x = O.normalize_batch_eval(O.get_patches(padded, pool, ps), stats).astype(np.float32)
    po, go = O.shrunk_class_gradients(layers, w, x)
    def worst(a, b):
        floor = 1e-3 * np.abs(b).max()
        return max(np.abs(a[:, :, t] - b[:, :, t]).max() / max(np.abs(b[:, :, t]).max(), floor) for t in range(b.shape[2]))
    print('tensor-core vs oracle %.3g, CUDA-core vs oracle %.3g, tensor-core vs CUDA-core %.3g' % (worst(g, go), worst(g2, go), worst(g, g2)))
    _assert_shrunk_close(g, g2)
    _assert_shrunk_close(g, go, outliers=0.02)
    _assert_shrunk_close(g2, go, outliers=0.02)


def _rand_A(n, tau, delta, seed, scale):
    rs = np.random.RandomState(seed)
    s = rs.randn(n, tau) * scale * np.exp(rs.randn(1, tau))
    p = rs.rand(n)
    g = np.stack([p[:, None] * s, -(1 - p)[:, None] * s])
    return O.gen_A_matrices(g[0], g[1], p, delta)


@pytest.mark.parametrize('n,tau,delta,scale', [(500, 7, 1e-5, 1e-2), (2000, 7, 1e-3, 1e-2), (300, 7, 1e-5, 1e-4),
                                              (40000, 3, 1e-3, 1e-1), (64, 16, 1e-2, 1.), (10, 1, 1e-3, 1.), (1, 4, 1e-3, 1.)])
def test_sdp_query_distribution(nb, n, tau, delta, scale):
    """Device solver == float64 restatement of the SDP: objective within 1e-3 (in fact within the certified tol), the
    certificate of the RETURNED q recomputed in float64, t = diag(M^-1), q on the simplex."""
    A = np.array(_rand_A(n, tau, delta, n + tau, scale))
    tol = 1e-4
    r = nb.get_engine().sdp_query_distribution(A, tol=tol)
    q = r['q']
    assert q.shape == (n,) and np.all(q >= 0) and abs(q.sum() - 1) < 1e-12
    phi, gap = O.sdp_certificate(A, q)
    assert abs(r['objective'] / phi - 1) < 1e-9 and abs(r['gap'] - gap) < 1e-6
    assert gap <= 2 * tol
    assert np.allclose(r['t'], np.diag(np.linalg.inv(np.tensordot(q, A, axes=(0, 0)))), rtol=1e-9)
    qo, to, phio, gapo, ito = O.sdp_solve(A, tol)
    assert abs(phi / phio - 1) < 3 * tol < OBJ_RTOL
    # shim with the cvxopt-shaped solution
    soln = nb.NNAL_tools.SDP_query_distribution(list(A), 0., None, 10)
    assert soln['status'] == 'optimal' and len(soln['x']) == n + tau
    assert np.allclose(np.array(soln['x'][:n]), q)
    assert abs(np.sum(soln['x'][n:]) / phi - 1) < 1e-9


def test_sdp_matches_slsqp_small(nb):
    """Independent solver (scipy SLSQP on the same programme) at small n."""
    A = np.array(_rand_A(40, 4, 1e-3, 5, 0.05))
    r = nb.get_engine().sdp_query_distribution(A, tol=1e-6)
    q2, phi2 = O.sdp_solve_slsqp(A)
    assert abs(r['objective'] / phi2 - 1) < 1e-5


def test_sdp_errors(nb):
    eng = nb.get_engine()
    with pytest.raises(Exception):
        eng.sdp_query_distribution(np.zeros((4, 17, 17)))
    with pytest.raises(Exception):
        eng.sdp_query_distribution(np.zeros((4, 3, 3)))           # singular: not positive definite


@pytest.mark.parametrize('B', [40, 10 ** 6])
def test_pw_fi_query_sdp_single(nb, B):
    """PW_NNAL.CNN_query(..., 'fi') with fi_mode='sdp' = the reference's own pipeline (PW_NNAL.py:89-163): same
    pre-filtered candidates, SDP objective within 1e-3 of the oracle's, and the sampled positions are exactly what
    sample_query_dstr returns for the device's q with the same uniform draws."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(120, 95)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    k = 8
    expr = Expr(k=k, B=B, lambda_=0., patch_shape=ps, ntb=128, stats=stats, fi_mode='sdp')
    np.random.seed(123)
    u = np.random.sample(k)
    np.random.seed(123)
    q, soln, sel = nb.fi.query_single_sdp(expr, model, None, padded, pool, return_solution=True)
    np.random.seed(123)
    q2 = nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'fi')
    assert np.array_equal(q, q2)
    qo, det = O.query_fi_sdp_single(layers, w, padded, pool, ps, 128, stats, k, B, u)
    assert set(sel.tolist()) == set(det['sel'].tolist())
    assert soln['status'] == 'optimal'
    assert abs(soln['primal objective'] / det['phi'] - 1) < OBJ_RTOL
    nB = len(sel)
    q_dev = np.array(soln['x'][:nB])
    # the device's q is near-optimal for the ORACLE's matrices too (candidate order may differ at exact ties only)
    pos = {int(v): i for i, v in enumerate(det['sel'])}
    perm = np.array([pos[int(v)] for v in sel])
    phi_o, gap_o = O.sdp_certificate(np.array(det['A'])[perm], q_dev)
    assert abs(phi_o / det['phi'] - 1) < OBJ_RTOL
    assert np.array_equal(q, sel[O.sample_query_dstr(q_dev.copy(), k, u)])
    assert len(q) <= k and len(np.unique(q)) == len(q) and np.all(np.isin(q, sel))


def test_pw_fi_query_sdp_multimg(nb):
    ps = (25, 25, 1)
    S, m = 3, 3
    shape = (34, 30, 4)
    allp, pools, st = [], [], np.zeros((S, 2 * m))
    rs = np.random.RandomState(44)
    for s in range(S):
        imgs = [np.clip(rs.randn(*shape) * 30 + 100, 0, None).astype(np.float32) for _ in range(m)]
        r = [(p - 1) // 2 for p in ps]
        allp.append([np.pad(im, ((r[0], r[0]), (r[1], r[1]), (r[2], r[2])), 'constant') for im in imgs] +
                    [(rs.rand(*shape) > .5).astype(np.int8)])
        pools.append(list(rs.choice(int(np.prod(shape)), [50, 0, 70][s], replace=False)))
        for j in range(m):
            st[s, 2 * j], st[s, 2 * j + 1] = imgs[j].mean(), imgs[j].std()
    layers = O.pw1_layers(2)
    w = O.he_init_weights(layers, (25, 25, 3), 62, bias_scale=0.05)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    k, B = 6, 30
    expr = Expr(k=k, B=B, lambda_=0., patch_shape=ps, ntb=128, fi_mode='sdp', SDP_solver='CVXOPT')
    expr.train_stats = st
    np.random.seed(5)
    u = np.random.sample(k)
    np.random.seed(5)
    Q, soln, G = nb.fi.query_multimg_sdp(expr, model, None, allp, pools, return_solution=True)
    np.random.seed(5)
    Q2 = nb.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'fi')
    assert len(Q) == S and all(np.array_equal(a, b) for a, b in zip(Q, Q2))
    assert len(Q[1]) == 0 and sum(len(a) for a in Q) <= k
    assert soln['status'] == 'optimal' and len(G) == B
    # the oracle's A-matrices of the same candidates (subject-major order, diag_load 1e-3, PW_NNAL.py:566-578)
    sizes = [len(p) for p in pools]
    local = O.global2local_inds(G, sizes)
    A = []
    for s in range(S):
        if len(local[s]) == 0:
            continue
        stats = [[st[s, 2 * j], st[s, 2 * j + 1]] for j in range(m)]
        x = O.normalize_batch_eval(O.get_patches(allp[s][:m], np.asarray(pools[s])[local[s]], ps), stats).astype(np.float32)
        po, go = O.shrunk_class_gradients(layers, w, x)
        A += O.gen_A_matrices(go[0], go[1], po[1], 1e-3)
    qo, to, phio, gapo, ito = O.sdp_solve(A, 1e-4)
    assert abs(soln['primal objective'] / phio - 1) < OBJ_RTOL
    q_dev = np.array(soln['x'][:B])
    draws = O.sample_query_dstr(q_dev.copy(), k, u)
    want = O.global2local_inds(G[draws], sizes)
    assert all(np.array_equal(np.sort(a), np.sort(b)) for a, b in zip(Q, want))


def test_whole_image_fi_query_sdp_multiclass(nb):
    """NNAL.CNN_query(..., 'fi') with fi_mode='sdp' on a c = 3 net (config-1 shape of the path): multiclass A-matrices
    (NNAL.py:354-414) from one backward pass per class, SDP, sampling."""
    rs = np.random.RandomState(21)
    x = rs.rand(300, 9, 7, 2).astype(np.float32)
    w = O.he_init_weights(SMALL, (9, 7, 2), 6, bias_scale=0.1)
    model = nb.NN.CNN((9, 7, 2), OrderedDict(SMALL), feature_layer=len(SMALL) - 2)
    model.set_weights(w)
    k, B = 10, 60
    expr = Expr(k=k, B=B, lambda_=0., batch_size=128, fi_mode='sdp')
    expr.pool_images = x
    np.random.seed(9)
    u = np.random.sample(k)
    np.random.seed(9)
    q, soln, sel = nb.fi.query_whole_sdp(model, expr, np.arange(300), None, return_solution=True)
    np.random.seed(9)
    q2 = nb.NNAL.CNN_query(model, expr, np.arange(300), 'fi', None)
    assert np.array_equal(q, q2) and soln['status'] == 'optimal'
    # oracle: same candidates (entropy pre-filter, NNAL_tools.uncertainty_filtering), A-matrices, SDP objective
    r = O.forward(SMALL, w, x)
    sel_o = O.uncertainty_filtering(r['posteriors'].copy(), B)
    assert set(sel.tolist()) == set(sel_o.tolist())
    po, go = O.shrunk_class_gradients(SMALL, w, x[sel])
    A = O.gen_A_matrices_multiclass(po.copy(), go)
    qo, to, phio, gapo, ito = O.sdp_solve(A, 1e-4)
    assert abs(soln['primal objective'] / phio - 1) < OBJ_RTOL
    q_dev = np.array(soln['x'][:B])
    phi_c, gap_c = O.sdp_certificate(A, q_dev)
    assert abs(phi_c / phio - 1) < OBJ_RTOL
    assert np.array_equal(q, sel[O.sample_query_dstr(q_dev.copy(), k, u)])


def test_shrunk_error_paths(nb):
    """Argument checks of the shrunk-gradient entry points: wrong patch shape -> ValueError (the reference's feed would
    fail on the placeholder shape), out-of-range voxel ids -> ValueError as np.unravel_index (patch_utils.py:1144)."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(8, 97)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_model(model, None)
    eng.upload(0, padded)
    st = np.array(stats, dtype=np.float64)
    with pytest.raises(ValueError):
        eng.fi_shrunk_voxels(0, pool, (23, 23, 1), st, shape=padded[0].shape)
    with pytest.raises(ValueError):
        eng.fi_shrunk_voxels(0, np.array([10 ** 9]), ps, st, shape=padded[0].shape)
    post, g = eng.fi_shrunk_voxels(0, pool[:0], ps, st, shape=padded[0].shape)
    assert post.shape == (2, 0) and g.shape == (2, 0, 7)
    # weights replaced -> the transposed fc planes of the backward pass are rebuilt (no stale gradients)
    p1, g1 = eng.fi_shrunk_voxels(0, pool, ps, st, shape=padded[0].shape)
    w2 = O.he_init_weights(layers, (25, 25, 3), 123, bias_scale=0.05)
    model.set_weights(w2)
    eng.set_model(model, None)
    p2, g2 = eng.fi_shrunk_voxels(0, pool, ps, st, shape=padded[0].shape)
    x = O.normalize_batch_eval(O.get_patches(padded, pool, ps), stats).astype(np.float32)
    po, go = O.shrunk_class_gradients(layers, w2, x)
    _assert_shrunk_close(g2, go)
    assert np.abs(g1 - g2).max() > 0


@pytest.mark.parametrize('opts', [
    {},
    {'bw_simt_fwd': 1, 'bw_no_tc': 1, 'bw_no_ws': 1, 'bw_no_tc8': 1, 'sdp_no_coop': 1},
    {'bw_no_ws': 1, 'bw_chunk': 16},
])
def test_fallback_kernels(nb, opts):
    """The fallback kernels behind the test-only switches (nnal_debug_option): fp32 CUDA-core forward, fp32 fc gradient,
    conv filter through L2, 4-channel register tile, one launch per SDP iteration -- same parity bar as the default kernels."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location('sdp_fallback_check', os.path.join(root, 'scripts', 'sdp_fallback_check.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    eng = nb.get_engine()
    for k, v in opts.items():
        eng.debug_option(k, v)
    try:
        mod.check()
    finally:
        for k in opts:
            eng.debug_option(k, 0)
    with pytest.raises(ValueError):
        eng.debug_option('no_such_switch', 1)


def test_sdp_from_shrunk_equals_host_assembly(nb):
    """nnal_sdp_from_shrunk assembles the A-matrices of gen_A_matrices on the device: same q (bit for bit -- same
    matrices, same deterministic solver) as the host assembly + nnal_sdp_query_distribution, incl. the clamped branches."""
    from nnal_b200.PW_NNAL import _A_from_shrunk
    rs = np.random.RandomState(31)
    n, tau = 3000, 7
    g = rs.randn(2, n, tau) * 1e-2
    p = rs.rand(n)
    p[:4] = [1e-9, 1 - 1e-9, 0., 1.]
    eng = nb.get_engine()
    A = _A_from_shrunk(g, p, 1e-5, as_list=False)
    assert all(np.array_equal(a, b) for a, b in zip(A, O.gen_A_matrices(g[0], g[1], p, 1e-5)))
    r0 = eng.sdp_query_distribution(A, tol=1e-4)
    r1 = eng.sdp_from_shrunk(g, p, 1e-5, tol=1e-4)
    assert r0['iterations'] == r1['iterations'] and np.array_equal(r0['q'], r1['q']) and r0['objective'] == r1['objective']
    soln = nb.NNAL_tools.SDP_query_distribution_from_shrunk(g, p, 1e-5, 10)
    assert soln['status'] == 'optimal' and np.array_equal(np.array(soln['x'][:n]), r0['q'])
    assert np.allclose(nb.NNAL_tools.solve_FIAL_SDP(list(A[:200])), eng.sdp_query_distribution(A[:200])['q'])


def test_config1_shape_whole_image_queries(nb):
    """BASELINE config 1 at reduced pool size: the PW1 layer dictionary on 28x28x1 inputs with 10 classes (He-normal
    weights, U[0,1) images) -- entropy, FI-trace and the literal multiclass FI pipeline through NNAL.CNN_query."""
    layers = O.pw1_layers(10)
    rs = np.random.RandomState(0)
    n = 160
    x = rs.rand(n, 28, 28, 1).astype(np.float32)
    w = O.he_init_weights(layers, (28, 28, 1), 1, bias_scale=0.0)
    model = nb.NN.CNN((28, 28, 1), OrderedDict(layers), feature_layer=len(layers) - 2)
    model.set_weights(w)
    r = O.forward(layers, w, x, feature_layer=len(layers) - 2)
    post = r['posteriors']
    # entropy: argsort(-H)[:10] (NNAL.py:298-310)
    expr = Expr(k=10, B=40, lambda_=0., batch_size=64)
    expr.pool_images = x
    q = nb.NNAL.CNN_query(model, expr, np.arange(n), 'entropy', None)
    H = O.compute_entropy(post.copy())
    kth = np.sort(-H)[9]
    assert len(q) == 10 and np.all(-H[q] <= kth + 1e-4) and np.all(np.isin(np.where(-H < kth - 1e-4)[0], q))
    # fi (default): top-10 of the last-layer FI trace (NNAL.py:121-139)
    q = nb.NNAL.CNN_query(model, expr, np.arange(n), 'fi', None)
    score = O.fi_trace_score(post, r['feature_layer'])
    kth = np.sort(-score)[9]
    tol = 1e-3 * np.abs(score).max()
    assert len(q) == 10 and np.all(-score[q] <= kth + tol)
    # fi_mode='sdp': multiclass A-matrices (ten backward passes), SDP, sampling
    expr.pars['fi_mode'] = 'sdp'
    np.random.seed(4)
    u = np.random.sample(10)
    np.random.seed(4)
    q, soln, sel = nb.fi.query_whole_sdp(model, expr, np.arange(n), None, return_solution=True)
    assert soln['status'] == 'optimal' and len(sel) == 40
    po, go = O.shrunk_class_gradients(layers, w, x[sel])
    eng = nb.get_engine()
    _, g = eng.fi_shrunk_images(x[sel])
    _assert_shrunk_close(g, go, outliers=0.05)         # zero biases + U[0,1) inputs: many activations at the ReLU edge
    A = O.gen_A_matrices_multiclass(po.copy(), go)
    qo, to, phio, gapo, ito = O.sdp_solve(A, 1e-4)
    assert abs(soln['primal objective'] / phio - 1) < OBJ_RTOL
    assert np.array_equal(q, sel[O.sample_query_dstr(np.array(soln['x'][:40]), 10, u)])


def test_sdp_device_solution_feasible_for_reference_programme(nb, golden):
    """The device solver's (q, t) satisfies the constraints the unmodified reference hands to cvxopt (golden) and reaches
    the reference objective c^T x of the float64 solution within the certified tolerance."""
    from tests.util import assert_feasible_for_reference_sdp
    r = nb.get_engine().sdp_query_distribution(golden['sdp_A'], tol=1e-6)
    obj = assert_feasible_for_reference_sdp(golden, r['q'], r['t'] * (1 + 1e-12))
    assert abs(obj / r['objective'] - 1) < 1e-9
    assert 0 <= obj / float(golden['sdp_phi']) - 1 < 1e-5


def _centred_features(d, n, seed):
    rs = np.random.RandomState(seed)
    X = np.maximum(rs.randn(d, n), 0)
    return X - X.mean(axis=1, keepdims=True)


@pytest.mark.parametrize('n,tau,d,lam', [(60, 5, 25, 2.0), (300, 7, 70, 0.05), (512, 7, 96, 0.5), (130, 3, 1, 1.0)])
def test_sdp_regularised_matches_oracle(nb, n, tau, d, lam):
    """lambda_ > 0 (NNAL_tools.py:625-644) on the device: every constraint holds to rounding, the objective equals the
    float64 oracle's (itself checked against SLSQP and the reference's captured programme) well inside 1e-3."""
    A = np.array(_rand_A(n, tau, 1e-3, n + d, 0.05))
    X = _centred_features(d, n, n)
    r = nb.get_engine().sdp_query_distribution_reg(A, lam, X, tol=1e-5)
    q = r['q']
    assert q.min() >= 0 and abs(q.sum() - 1) < 1e-10
    assert np.abs(X @ q).max() < 1e-9 * np.abs(X).max()
    M = np.tensordot(q, A, axes=(0, 0))
    Phi = np.trace(np.linalg.inv(M)) - lam * (np.sum(X ** 2, axis=0) @ q)
    assert abs(r['objective'] - Phi) < 1e-9 * abs(Phi)            # the reported objective is the objective of the returned q
    assert np.allclose(r['t'], np.diag(np.linalg.inv(M)), rtol=1e-9)
    qo, to, Phio, gapo, ito = O.sdp_solve_reg(A, lam, X, 1e-6)
    assert r['gap'] <= 1e-5 and Phi >= Phio - 2e-6 * abs(Phio)    # nobody beats the (tighter) oracle optimum ...
    assert abs(Phi / Phio - 1) < 1e-4                             # ... and the device is within its certificate of it
    soln = nb.NNAL_tools.SDP_query_distribution(list(A), lam, X, 10, tol=1e-5)
    assert soln['status'] == 'optimal' and np.allclose(np.array(soln['x'][:n]), q)


def test_sdp_regularised_small_vs_slsqp_and_reference_programme(nb, golden):
    """Independent solver at small n, and feasibility for the constraint data the UNMODIFIED reference builds for
    lambda_ > 0 (c, A_eq, b_eq captured by oracle/check_against_reference.py)."""
    A, X, lam = golden['sdp_A'], golden['sdpr_X'], float(golden['sdpr_lambda'])
    r = nb.get_engine().sdp_query_distribution_reg(A, lam, X, tol=1e-7)
    x = np.concatenate([r['q'], r['t']])
    assert np.abs(golden['sdpr_Aeq'] @ x - golden['sdpr_beq'][:, 0]).max() < 1e-10       # X q = 0, sum q = 1
    assert abs((golden['sdpr_c'][:, 0] @ x).item() - r['objective']) < 1e-9 * abs(r['objective'])
    assert abs(r['objective'] / float(golden['sdpr_Phi']) - 1) < 1e-5
    q2, Phi2 = O.sdp_solve_reg_slsqp(A, lam, X)
    assert abs(r['objective'] / Phi2 - 1) < 1e-5
    assert_feasible_lmis(golden, r['q'], r['t'])


def assert_feasible_lmis(golden, q, t):
    """The LMIs are those of the lambda_ = 0 programme (inequality_cvx_matrix does not depend on lambda_)."""
    x = np.concatenate([q, t])
    tau = len(t)
    for j in range(tau + 1):
        G, h = golden['sdp_G%d' % j], golden['sdp_h%d' % j]
        m = h.shape[0]
        slack = h - (G @ x).reshape(m, m)
        slack = (slack + slack.T) / 2
        assert np.linalg.eigvalsh(slack).min() > -1e-9 * np.abs(slack).max(), j


def test_pw_fi_query_sdp_single_regularised_dispatch(nb, golden):
    """PW_NNAL.CNN_query 'fi' with lambda_ > 0 on the inputs of the golden run of the unmodified reference: candidates'
    features -> refine_feature_matrix -> zero-mean rows -> regularised programme -> sampler."""
    from collections import OrderedDict
    from tests.test_reference_dispatch_golden import LAYERS, M, PS, _case
    w, allp, pools, st, stats0 = _case(golden)
    pool0 = np.array(pools[0])
    model = nb.NN.CNN((5, 5, M), OrderedDict(LAYERS), feature_layer=len(LAYERS) - 2)
    model.set_weights(w)
    expr = Expr(k=9, B=30, lambda_=0.2, patch_shape=PS, ntb=16, stats=stats0, fi_mode='sdp')
    np.random.seed(77)
    qf, soln, sel = nb.fi.query_single_sdp(expr, model, None, allp[0][:M], pool0, return_solution=True)
    _, det = O.query_fi_sdp_single(LAYERS, w, allp[0][:M], pool0, PS, 16, stats0, 9, 30, golden['q_fi_u'], diag_load=1e-5, lambda_=0.2)
    assert soln['status'] == 'optimal'
    assert abs(soln['primal objective'] / det['phi'] - 1) < 1e-3
    q_dev = np.array(soln['x'][:len(sel)])
    assert np.abs(det['ref_F'] @ q_dev).max() < 1e-6 * np.abs(det['ref_F']).max()      # the oracle's constraints hold for the device's q
    assert np.array_equal(qf, sel[O.sample_query_dstr(q_dev.copy(), 9, golden['q_fi_u'])])
    got, want = set(np.asarray(qf).tolist()), set(golden['q_fi_sdp_single_lambda'].tolist())
    assert got <= set(sel.tolist()) and len(got & want) >= len(want) - 3, (sorted(got), sorted(want))
