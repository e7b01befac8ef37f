"""GPU parity tests of the Fisher-information path (``pytest -m gpu`` on the B200 box): factored
last-layer FI, greedy selection, weighted Gram on tensor cores, the 'fi' query branches -- all through
the C ABI -- against oracle/fi_oracle.py (float64)."""
import numpy as np
import pytest

import oracle as O
from tests.util import centered_weights, pad_imgs, synth_volume, vol_stats

pytestmark = pytest.mark.gpu

OBJ_RTOL = 1e-3      # north_star: FI objectives within 1e-3 relative


class Expr(object):
    def __init__(self, **pars):
        self.pars = pars
        self.nclass = 2


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


def _factors(n, d, dp, seed):
    rs = np.random.RandomState(seed)
    U = np.maximum(rs.randn(n, d), 0).astype(np.float32) * 2
    A = np.maximum(rs.randn(n, dp), 0).astype(np.float32) * 2
    Wl = (rs.randn(2, d) * .3).astype(np.float32)
    p1 = rs.rand(n)
    p1[:3] = [1e-9, 1 - 1e-9, 0.5]
    return p1, U, A, Wl


def _oracle_kernel(p1, U, A, Wl, two):
    Kt = O.last_layers_kernel(p1, U.T.astype(np.float64), A.T.astype(np.float64) if two else None,
                              Wl.astype(np.float64) if two else None)
    D = O.last_layers_dim(2, U.shape[1], A.shape[1] if two else None)
    return Kt, D


def reduced_objectives(Kt, delta, sel):
    """tr((delta I + Kt_SS/s)^-1) for every prefix S of ``sel`` -- the kernel-dependent part of the objective (the
    rest, (D - s)/delta, is a constant ~1e12 that would hide any error at rtol 1e-3)."""
    out = []
    for s in range(1, len(sel) + 1):
        S = list(sel[:s])
        out.append(np.trace(np.linalg.inv(delta * np.eye(s) + Kt[np.ix_(S, S)] / float(s))))
    return np.array(out)


def assert_greedy_equivalent(Kt, D, delta, sel, red, rtol=OBJ_RTOL, red_oracle=None):
    """``sel`` must be a valid greedy trajectory up to ties within ``rtol``; ``red`` the REDUCED objective
    tr((delta I + K_SS/s)^-1) reported for it per step, which must equal the float64 value for that same selection and,
    at the last step, the oracle's own greedy value (``red_oracle``)."""
    sel = np.asarray(sel)
    assert len(np.unique(sel)) == len(sel)
    rep = O.greedy_fi_replay(Kt, D, delta, sel)
    assert np.all(rep[:, 0] <= rep[:, 1] * (1 + rtol) + 1e-300), 'picked a candidate clearly worse than the best'
    want = reduced_objectives(Kt, delta, sel)
    assert np.allclose(red, want, rtol=rtol), 'reduced objective off by %g' % np.abs(red / want - 1).max()
    if red_oracle is not None:
        assert abs(red[-1] / red_oracle[len(sel) - 1] - 1) < rtol, 'final reduced objective differs from the oracle greedy'


@pytest.mark.parametrize('two', [False, True])
@pytest.mark.parametrize('n,d,dp,k', [(300, 64, 48, 40), (1000, 256, 128, 25), (37, 20, 12, 37), (500, 30, 18, 170)])
def test_fi_greedy_given_factors(nb, two, n, d, dp, k):
    """Greedy selection on host-provided factors == the float64 oracle (same float32 inputs)."""
    p1, U, A, Wl = _factors(n, d, dp, n + d)
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U, A if two else None, Wl if two else None)
    info = eng.fi_info()
    Kt, D = _oracle_kernel(p1, U, A, Wl, two)
    assert info['n'] == n and info['n_layers'] == (2 if two else 1) and info['D'] == D
    delta = 1e-3
    sel, obj, red = eng.fi_greedy(k, delta)
    So, oo, ro = O.greedy_fi_rank1(Kt, D, delta, k, return_reduced=True)
    # kernel entries carry ~1e-8 relative error (float32 FMA chains flushed into float64 sums): the kernel-dependent
    # part of the objective is reproduced to ~1e-5, far inside the 1e-3 of the north star
    assert_greedy_equivalent(Kt, D, delta, sel, red, rtol=1e-4, red_oracle=ro)
    assert len(set(sel.tolist()) ^ set(So.tolist())) <= 2
    assert np.allclose(obj, (D - np.arange(1, len(sel) + 1)) / delta + red, rtol=1e-12)


def test_fi_greedy_large_candidate_set_f16_rows(nb):
    """From 32,768 candidates on, the kernel-column pass reads compact fp16 copies of the candidates' factor rows
    (csrc/fi.cu F16_MIN).  Held to the north star's bar for FI objectives (1e-3 relative on the kernel-dependent part) against the
    float64 oracle on the float32 factors: every pick within 1e-3 of the best available candidate, the reported reduced objective
    within 1e-3 of the float64 value of the selected set and of the oracle's own greedy, (nearly) the same selected set."""
    n, d, dp, k = 40000, 64, 32, 24
    p1, U, A, Wl = _factors(n, d, dp, 5)
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U, A, Wl)
    delta = 1e-3
    sel, obj, red = eng.fi_greedy(k, delta)
    # oracle on the candidates that matter (the full 40000^2 kernel would be 12.8 GB): the device's picks + the oracle's picks
    # are found inside the 4000 candidates with the largest diagonal... not guaranteed -> use the exact column-wise oracle
    Ut, At = U.T.astype(np.float64), A.T.astype(np.float64)
    w = p1 * (1 - p1)
    sw = np.sqrt(w)
    beta2 = (Wl[0].astype(np.float64) - Wl[1].astype(np.float64)) ** 2

    def column(j):
        uu = Ut.T @ Ut[:, j]
        aa = At.T @ At[:, j]
        mm = ((Ut > 0).T * beta2) @ (Ut[:, j] > 0)
        return sw * sw[j] * (2. * (uu + 1.) + mm * (aa + 1.))
    diag = np.array([sw[j] ** 2 * (2. * (Ut[:, j] @ Ut[:, j] + 1.) + ((Ut[:, j] > 0) @ beta2) * (At[:, j] @ At[:, j] + 1.)) for j in range(n)])
    avail = np.ones(n, dtype=bool)
    cols, So, ro = [], [], []
    for t in range(k):
        alpha = (t + 1) * delta
        if t == 0:
            rr, e, trC = diag.copy(), np.zeros(n), 0.
        else:
            kj = np.stack(cols, axis=1)
            C = np.linalg.inv(alpha * np.eye(t) + kj[So])
            Y = kj @ C
            rr, e, trC = diag - np.sum(Y * kj, axis=1), np.sum(Y * Y, axis=1), np.trace(C)
        loss = (1. + e) / (alpha + rr)
        loss[~avail] = np.inf
        # the device's pick must be within 1e-3 of the best available loss of the ORACLE's trajectory while they coincide
        j = int(np.argmin(loss))
        if list(sel[:t]) == So:
            assert loss[sel[t]] <= loss[j] * (1 + OBJ_RTOL)
        So.append(j)
        avail[j] = False
        ro.append((t + 1) * (trC + loss[j]))
        cols.append(column(j))
    assert len(set(sel.tolist()) ^ set(So)) <= 2
    Ks = np.stack([column(j)[sel] for j in sel], axis=1)
    want = np.array([np.trace(np.linalg.inv(delta * np.eye(s) + Ks[:s, :s] / float(s))) for s in range(1, k + 1)])
    assert np.allclose(red, want, rtol=OBJ_RTOL), np.abs(red / want - 1).max()
    assert abs(red[-1] / ro[-1] - 1) < OBJ_RTOL


@pytest.mark.parametrize('n,d,dp,k,two', [(3000, 256, 128, 40, True), (700, 64, 32, 128, True), (900, 128, 64, 33, False)])
def test_fi_greedy_pipelined_and_speculative_equal_plain(nb, n, d, dp, k, two):
    """The single-process greedy runs a pipelined step (shifted inverse one step ahead + bordering) with speculative
    kernel columns (csrc/fi.cu).  Neither may change the result: the same selection, in the same order, as the plain step
    (debug flag 8) and as the pipelined step without speculation (flag 16); reduced objectives equal to 1e-9 (the inverse
    is reached by a different, equally stable, elimination order)."""
    p1, U, A, Wl = _factors(n, d, dp, 3 * n + k)
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U, A if two else None, Wl if two else None)
    out = {}
    try:
        for flag in (0, 16, 8):
            eng.debug_option('fi_flags', flag)
            out[flag] = eng.fi_greedy(k, 1e-3)
    finally:
        eng.debug_option('fi_flags', 0)
    for flag in (16, 8):
        assert np.array_equal(out[0][0], out[flag][0]), 'selection differs (flag %d)' % flag
        assert np.allclose(out[0][2], out[flag][2], rtol=1e-9), np.abs(out[0][2] / out[flag][2] - 1).max()
    Kt, D = _oracle_kernel(p1, U, A, Wl, two)
    assert_greedy_equivalent(Kt, D, 1e-3, out[0][0], out[0][2], rtol=1e-4)


def test_fi_greedy_edge_cases(nb):
    eng = nb.get_engine()
    p1, U, A, Wl = _factors(5, 16, 8, 1)
    eng.fi_set_factors(p1, U, A, Wl)
    sel, obj, _ = eng.fi_greedy(50, 1e-3)                 # k > n: every candidate, once
    assert sorted(sel.tolist()) == list(range(5))
    eng.fi_set_factors(p1[:1], U[:1], A[:1], Wl)
    sel, obj, _ = eng.fi_greedy(1, 1e-3)
    assert sel.tolist() == [0]
    # duplicated candidates: exact ties -> lowest index
    U2 = np.concatenate([U, U]); A2 = np.concatenate([A, A]); p2 = np.concatenate([p1, p1])
    eng.fi_set_factors(p2, U2, A2, Wl)
    sel, _, _ = eng.fi_greedy(3, 1e-3)
    Kt, D = _oracle_kernel(p2, U2, A2, Wl, True)
    assert np.array_equal(sel, O.greedy_fi_rank1(Kt, D, 1e-3, 3)[0])
    with pytest.raises(ValueError):
        eng.fi_set_factors(p1, U, A, Wl[:, :5])


def test_fi_step_protocol_two_contexts(nb):
    """The multi-GPU step protocol (local best -> global arg-min -> winner factors -> apply) run over two
    contexts that split the candidates reproduces the single-context greedy."""
    n, d, dp, k = 400, 96, 64, 30
    p1, U, A, Wl = _factors(n, d, dp, 9)
    delta = 1e-3
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U, A, Wl)
    ref_sel, ref_obj, _ = eng.fi_greedy(k, delta)
    cut = 170
    parts = [(0, cut), (cut, n)]
    engs = [nb.Engine(0), nb.Engine(0)]
    try:
        for e, (a, b) in zip(engs, parts):
            e.fi_set_factors(p1[a:b], U[a:b], A[a:b], Wl)
            e.fi_begin(k, delta)
        D = engs[0].fi_info()['D']
        sel, obj = [], []
        for t in range(k):
            best = []
            for r, (e, (a, b)) in enumerate(zip(engs, parts)):
                loss, cand, trc = e.fi_step_local_best(t)
                best.append((loss, cand + a if cand >= 0 else 1 << 60, r, cand, trc))
            loss, gid, owner, cand, trc = min(best)
            f = engs[owner].fi_winner_factors(t, cand)
            for r, e in enumerate(engs):
                e.fi_step_apply(t, f, r == owner, cand if r == owner else 0)
            sel.append(gid)
            obj.append((D - (t + 1)) / delta + (t + 1) * (trc + loss))
    finally:
        for e in engs:
            e.close()
    assert sel == ref_sel.tolist()
    assert np.allclose(obj, ref_obj, rtol=1e-9)


@pytest.mark.parametrize('n,d', [(50, 127), (700, 64), (3000, 200)])
def test_fi_gram_tensor_cores(nb, n, d):
    """H = sum_i wq_i [u_i;1][u_i;1]^T on the tcgen05 GEMM vs float64; primal objective through H."""
    p1, U, A, Wl = _factors(n, d, 8, n)
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U)
    rs = np.random.RandomState(0)
    for q in (None, rs.dirichlet(np.ones(n))):
        H = eng.fi_gram(q)
        wq = (np.full(n, 1. / n) if q is None else q) * p1 * (1 - p1)
        Ho = O.weighted_gram(U.T.astype(np.float64), wq)
        assert H.shape == (d + 1, d + 1)
        assert np.abs(H - Ho).max() < 1e-5 * np.abs(Ho).max()
        delta = 1e-2
        f_gpu = O.fi_objective_from_gram(H.astype(np.float64), 2, delta)
        f_ora = O.fi_objective_from_gram(Ho, 2, delta)
        assert abs(f_gpu / f_ora - 1) < OBJ_RTOL
    assert np.array_equal(eng.fi_gram_read(), H)
    ptr, rows, ld = eng.fi_gram_device()
    assert ptr and rows == d + 1 and ld >= rows


def test_fi_gram_equals_dual_objective(nb):
    """Primal (Gram) and dual (kernel) forms of tr((sum_i q_i A_i)^-1) agree at q = uniform(S)."""
    n, d = 40, 24
    p1, U, A, Wl = _factors(n, d, 8, 77)
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U)
    delta = 1e-2
    sel, obj, _ = eng.fi_greedy(12, delta)
    q = np.zeros(n)
    q[sel] = 1. / len(sel)
    H = eng.fi_gram(q).astype(np.float64)
    assert abs(O.fi_objective_from_gram(H, 2, delta) / obj[-1] - 1) < OBJ_RTOL


@pytest.mark.parametrize('n,d,k', [(300, 127, 20), (900, 320, 64)])
def test_fi_gram_subset_and_solve(nb, n, d, k):
    """Gram over the support of a query distribution + the float64 Gauss-Jordan primal objective on the device:
    tr((delta I + 2 H_S)^-1) against NumPy, and the Fisher-information ratio against a second (pool-wide) Gram."""
    import torch
    p1, U, A, Wl = _factors(n, d, 8, n + 1)
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U)
    rs = np.random.RandomState(3)
    S = rs.choice(n, k, replace=False)
    w = p1 * (1 - p1)
    Hp = eng.fi_gram(None)                                       # pool-wide Gram, q = 1/n
    ptr, rows, ld = eng.fi_gram_device()
    from nnal_b200 import dist
    G2 = dist.device_view(ptr, (rows * ld,), '<f4').clone()
    torch.cuda.synchronize()
    Hs = eng.fi_gram_subset(S, np.full(k, 1. / k), read=True)
    Ut = np.concatenate([U.astype(np.float64), np.ones((n, 1))], axis=1)
    Hso = (Ut[S] * (w[S] / k)[:, None]).T @ Ut[S]
    assert np.abs(Hs - Hso).max() < 1e-5 * np.abs(Hso).max()
    for delta in (1e-2, 1e-3):
        tr, ratio = eng.fi_gram_solve(delta, 2.0, G2.data_ptr())
        Minv = np.linalg.inv(delta * np.eye(d + 1) + 2. * Hs.astype(np.float64))
        assert abs(tr / np.trace(Minv) - 1) < 1e-9               # same float32 Gram in, float64 arithmetic
        want = np.sum(Minv * (delta * np.eye(d + 1) + 2. * Hp.astype(np.float64)))
        assert abs(ratio / want - 1) < 1e-8
        # against the dual (kernel) form of the same objective
        Kss = 2. * (Ut[S] @ Ut[S].T) * np.sqrt(np.outer(w[S], w[S]))
        dual = (d + 1 - k) / delta + np.trace(np.linalg.inv(delta * np.eye(k) + Kss / k))
        assert abs(tr / dual - 1) < OBJ_RTOL
    # empty subset: H = 0
    H0 = eng.fi_gram_subset(np.zeros(0, dtype=np.int64), np.zeros(0), read=True)
    assert not H0.any()
    tr, _ = eng.fi_gram_solve(1e-2)
    assert abs(tr / ((d + 1) / 1e-2) - 1) < 1e-12


def test_pw_fi_report_primal_equals_dual(nb):
    """fi_report: the selection's Gram (tensor cores) -> primal objective == the dual objective of the greedy loop
    (last-layer FI, multi-volume diag_load 1e-3)."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(260, 90)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    expr = Expr(k=10, B=80, lambda_=0., patch_shape=ps, ntb=128, stats=stats, fi_layers=1, fi_diag_load=1e-3, fi_report=True)
    q, obj, red = nb.fi.query_single(expr, model, None, padded, pool, return_objective='reduced')
    rep = nb.fi.last_report
    assert rep['k'] == 10 and rep['n_candidates'] == 80
    # the Gram is float32 (split-fp16 tensor-core products, fp32 accumulate): at diag_load 1e-3 the (d+1)^2 inverse
    # reproduces the objective to ~1e-4 (measured 6e-5); the kernel-dependent part, ~1e-6 of the trace, is below that
    # resolution and is checked through the float64 dual form instead
    assert abs(rep['primal_last_layer'] / obj[-1] - 1) < OBJ_RTOL
    assert abs(rep['dual_last_layer'] / obj[-1] - 1) < 1e-9
    assert abs(rep['dual_reduced'] / red[-1] - 1) < 1e-4
    assert rep['fi_ratio'] > 0


def _pw_setup(n_pool, seed, shape=(40, 36, 6)):
    ps = (25, 25, 1)
    imgs = synth_volume(shape, 3, seed)
    padded = pad_imgs(imgs, ps)
    stats = vol_stats(imgs)
    rs = np.random.RandomState(seed + 1)
    pool = rs.choice(int(np.prod(shape)), n_pool, replace=False).astype(np.int64)
    layers = O.pw1_layers(2)
    probe = O.normalize_batch_eval(O.get_patches(padded, pool[:64], ps), stats).astype(np.float32)
    w = centered_weights(layers, (25, 25, 3), seed + 2, probe)
    return ps, imgs, padded, stats, pool, layers, w


@pytest.mark.parametrize('nl,B', [(2, 60), (1, 60), (2, 10 ** 6)])
def test_pw_fi_query_single(nb, nl, B):
    """PW_NNAL.CNN_query(..., 'fi') on PW1: selection follows the oracle's greedy criterion (ties within
    1e-3) and the FI objective matches within 1e-3 relative."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(220, 70)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    k = 12
    expr = Expr(k=k, B=B, lambda_=0., patch_shape=ps, ntb=128, stats=stats, fi_layers=nl, fi_diag_load=1e-5)
    q, obj, red = nb.fi.query_single(expr, model, None, padded, pool, return_objective='reduced')
    q2 = nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'fi')
    assert np.array_equal(q, q2)
    qo, oo, det = O.query_fi_single(layers, w, padded, pool, ps, 128, stats, k, B, nl, 1e-5)
    # same candidate set (pre-filter) up to posterior tolerance, then greedy equivalence on the oracle kernel
    sel = det['sel']
    assert np.all(np.isin(q, sel))
    pos = {int(v): i for i, v in enumerate(sel)}
    S_gpu = np.array([pos[int(v)] for v in q])
    ro = reduced_objectives(det['Kt'], 1e-5, [pos[int(v)] for v in qo])          # the oracle's own greedy trajectory
    assert_greedy_equivalent(det['Kt'], det['D'], 1e-5, S_gpu, red, red_oracle=ro)


def test_pw_fi_query_multimg(nb):
    ps = (25, 25, 1)
    S, m = 3, 3
    shape = (34, 30, 4)
    allp, pools, st = [], [], np.zeros((S, 2 * m))
    rs = np.random.RandomState(43)
    for s in range(S):
        imgs = synth_volume(shape, m, 80 + s)
        allp.append(pad_imgs(imgs, ps) + [(rs.rand(*shape) > .5).astype(np.int8)])
        pools.append(list(rs.choice(int(np.prod(shape)), [90, 0, 130][s], replace=False)))
        for j in range(m):
            st[s, 2 * j], st[s, 2 * j + 1] = imgs[j].mean(), imgs[j].std()
    layers = O.pw1_layers(2)
    probe = O.normalize_batch_eval(O.get_patches(allp[0][:m], pools[0][:64], ps),
                                   [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(m)]).astype(np.float32)
    w = centered_weights(layers, (25, 25, 3), 61, probe)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    k, B = 8, 50
    expr = Expr(k=k, B=B, lambda_=0., patch_shape=ps, ntb=128, SDP_solver='CVXOPT')
    expr.train_stats = st
    Q, obj, red = nb.fi.query_multimg(expr, model, None, allp, pools, return_objective='reduced')
    Q2 = nb.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'fi')
    assert len(Q) == S and len(Q[1]) == 0 and sum(len(a) for a in Q) == k
    assert all(np.array_equal(a, b) for a, b in zip(Q, Q2))
    Qo, oo, det = O.query_fi_multimg(layers, w, allp, pools, ps, 128, st, k, B, 2, 1e-3)
    # map the GPU's picks to the oracle's candidate order (subject-major) and replay
    sel_inds = det['sel_inds']
    offs = np.concatenate([[0], np.cumsum([len(x) for x in sel_inds])])
    S_gpu = []
    for s in range(S):
        lookup = {int(v): i for i, v in enumerate(sel_inds[s])}
        for v in Q[s]:
            assert int(v) in lookup, 'picked a sample outside the pre-filtered candidates'
    # greedy order is lost by the per-subject split; check the objective of the final set instead
    cand = np.concatenate([[offs[s] + {int(v): i for i, v in enumerate(sel_inds[s])}[int(v)] for v in Q[s]]
                           for s in range(S)]).astype(int)
    Kss = det['Kt'][np.ix_(cand, cand)]
    red_set = np.trace(np.linalg.inv(1e-3 * np.eye(k) + Kss / float(k)))          # reduced objective of the GPU's set, float64
    assert abs(red_set / red[-1] - 1) < OBJ_RTOL
    cand_o = np.concatenate([[offs[s] + {int(v): i for i, v in enumerate(sel_inds[s])}[int(v)] for v in Qo[s]]
                             for s in range(S)]).astype(int)
    Kso = det['Kt'][np.ix_(cand_o, cand_o)]
    assert abs(red[-1] / np.trace(np.linalg.inv(1e-3 * np.eye(k) + Kso / float(k))) - 1) < OBJ_RTOL
    assert abs(obj[-1] - ((det['D'] - k) / 1e-3 + red[-1])) <= 1e-12 * obj[-1]


def test_whole_image_fi_trace_score(nb):
    """NNAL.CNN_query(..., 'fi') for c > 2: top-k of the last-layer FI trace (NNAL.py:121-139)."""
    from collections import OrderedDict
    from tests.test_gpu_parity import SMALL
    rs = np.random.RandomState(12)
    x = rs.rand(400, 9, 7, 2).astype(np.float32)
    w = O.he_init_weights(SMALL, (9, 7, 2), 5, bias_scale=0.1)
    model = nb.NN.CNN((9, 7, 2), OrderedDict(SMALL), feature_layer=len(SMALL) - 2)
    model.set_weights(w)
    expr = Expr(k=20, B=100, lambda_=0., batch_size=128)
    expr.pool_images = x
    q = nb.NNAL.CNN_query(model, expr, np.arange(400), 'fi', None)
    r = O.forward(SMALL, w, x, feature_layer=len(SMALL) - 2)
    score = O.fi_trace_score(r['posteriors'], r['feature_layer'])
    kth = np.sort(-score)[19]
    tol = 1e-3 * np.abs(score).max()
    assert len(q) == 20 and np.all(-score[q] <= kth + tol)
    assert np.all(np.isin(np.where(-score < kth - tol)[0], q))


def test_fi_device_message_protocol_two_contexts(nb):
    """Device-resident step protocol (pack -> all-gather -> apply) over two contexts that split the candidates,
    the all-gather emulated by device copies, reproduces the single-context greedy."""
    import torch
    n, d, dp, k = 500, 128, 64, 40
    p1, U, A, Wl = _factors(n, d, dp, 19)
    delta = 1e-3
    eng = nb.get_engine()
    eng.fi_set_factors(p1, U, A, Wl)
    ref_sel, ref_obj, ref_red = eng.fi_greedy(k, delta)
    parts = [(0, 210), (210, 210), (210, n)]              # middle rank owns no candidates
    engs = [nb.Engine(0) for _ in parts]
    try:
        for e, (a, b) in zip(engs, parts):
            e.fi_set_factors(p1[a:b], U[a:b], A[a:b], Wl)
            e.fi_begin(k, delta)
            e.fi_set_gids(np.arange(a, b))
        nbytes = engs[0].fi_msg_bytes()
        world = len(engs)
        send = [torch.zeros(nbytes, dtype=torch.uint8, device='cuda') for _ in engs]
        recv = torch.zeros(world * nbytes, dtype=torch.uint8, device='cuda')
        for t in range(k):
            for r, e in enumerate(engs):
                e.fi_step_pack(t, send[r].data_ptr())
                e.synchronize()
            for r in range(world):
                recv[r * nbytes:(r + 1) * nbytes].copy_(send[r])
            torch.cuda.synchronize()
            for r, e in enumerate(engs):
                e.fi_step_apply_gathered(t, recv.data_ptr(), world, r)
                e.synchronize()
        res = [e.fi_result(k) for e in engs]
    finally:
        for e in engs:
            e.close()
    for sel, red in res:
        assert np.array_equal(sel, ref_sel)
        assert np.allclose(red, ref_red, rtol=1e-12)
