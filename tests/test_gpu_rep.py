"""GPU parity tests of the representativeness queries ('rep-entropy', 'core-set') against oracle/rep_oracle.py."""
from collections import OrderedDict

import numpy as np
import pytest

import oracle as O
from tests.util import centered_weights, pad_imgs, synth_volume

pytestmark = pytest.mark.gpu

NET = [('conv1', [6, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']), ('conv2', [8, 'conv', [3, 3]]), ('max2', [[2, 2], 'pool']),
       ('fc1', [40, 'fc']), ('fc2', [24, 'fc']), ('fc3', [2, 'fc'])]


class Expr(object):
    def __init__(self, **pars):
        self.pars = pars


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


def _whole(nb, n, seed):
    rs = np.random.RandomState(seed)
    x = rs.rand(n, 9, 7, 2).astype(np.float32)
    w = O.he_init_weights(NET, (9, 7, 2), seed + 1, bias_scale=0.1)
    model = nb.NN.CNN((9, 7, 2), OrderedDict(NET), feature_layer=len(NET) - 2)
    model.set_weights(w)
    return x, w, model


def test_rep_entropy_whole_image(nb):
    x, w, model = _whole(nb, 600, 5)
    expr = Expr(k=15, B=80, lambda_=0., batch_size=128)
    expr.pool_images = x
    q = nb.NNAL.CNN_query(model, expr, np.arange(600), 'rep-entropy', None)
    qo, det = O.query_rep_entropy_whole(NET, w, x, 15, 80)
    assert len(q) == 15 and len(np.unique(q)) == 15 and np.all(np.isin(q, det['sel']))
    pos = {int(v): i for i, v in enumerate(det['sel'])}
    rep = O.facility_location_replay(det['sims'], [pos[int(v)] for v in q])
    assert np.all(rep[:, 0] >= rep[:, 1] - 1e-4 * np.abs(rep[:, 1])), 'picked a column clearly worse than the best'
    assert len(set(q.tolist()) ^ set(qo.tolist())) <= 2


def test_cross_sims_and_kcenter_given_pool(nb):
    """get_cross_sims-style row maxima and the k-center loop on the pool features of a forward pass."""
    x, w, model = _whole(nb, 500, 8)
    eng = nb.get_engine()
    eng.set_model(model)
    eng.pool_begin(500, 1)
    eng.pool_eval_images(x, 0)
    F = eng.pool_features().astype(np.float64)                 # [d, n]
    T = F[:, ::37][:, :9]
    sims = eng.cross_sims(np.ascontiguousarray(T.T))
    ref = O.get_cross_sims(F, T)
    assert np.allclose(sims, ref, rtol=0, atol=2e-6)
    eng.cs_begin(2, None, None, 20)
    sel, val = eng.cs_greedy(20)
    rep = O.kcenter_replay(F, ref, sel)
    assert len(np.unique(sel)) == 20
    assert np.all(rep[:, 0] <= rep[:, 1] + 1e-5), 'picked a sample clearly more similar than the least similar one'
    assert np.allclose(val, rep[:, 0], atol=1e-5)
    # no labeled set: start from -inf -> first pick is the lowest index
    eng.cs_begin(0, None, None, 3)
    sel0, _ = eng.cs_greedy(3)
    assert sel0[0] == 0


def test_rep_and_kcenter_two_contexts(nb):
    """Sharded protocols (score all-reduce for facility location, message all-gather for k-center) emulated with
    two contexts that split the pool rows: same selections as one context."""
    import torch
    x, w, model = _whole(nb, 400, 11)
    eng = nb.get_engine()
    eng.set_model(model)
    eng.pool_begin(400, 1)
    eng.pool_eval_images(x, 0)
    post = eng.pool_posteriors().astype(np.float64)
    H = O.compute_entropy(post.copy())
    sel = O.stable_topk(-H, 50)
    cols = eng.pool_feature_rows(sel)
    eng.rep_set(cols, sel, 12)
    ref_sel, ref_val = eng.rep_greedy(12)
    Fall = eng.pool_features().astype(np.float64)
    T = np.ascontiguousarray(Fall[:, 5:12].T)
    eng.cross_sims(T)
    eng.cs_begin(2, None, None, 10)
    ref_cs, ref_csv = eng.cs_greedy(10)
    cut = 170
    parts = [(0, cut), (cut, 400)]
    engs = [nb.Engine(0), nb.Engine(0)]
    try:
        for e, (a, b) in zip(engs, parts):
            e.set_model(model)
            e.pool_begin(b - a, 1)
            e.pool_eval_images(x[a:b], 0)
            excl = sel[(sel >= a) & (sel < b)] - a
            e.rep_set(cols, excl, 12)
        sc = [torch.zeros(50, dtype=torch.float64, device='cuda') for _ in engs]
        for t in range(12):
            for e, s_ in zip(engs, sc):
                e.rep_step_scores(s_.data_ptr())
                e.synchronize()
            tot = sc[0] + sc[1]
            torch.cuda.synchronize()
            for e in engs:
                e.rep_step_pick(t, tot.data_ptr())
                e.synchronize()
        for e in engs:
            s2, v2 = e.sel_result(12)
            assert np.array_equal(s2, ref_sel)
            assert np.allclose(v2, ref_val, rtol=1e-6)        # float32 row runs are cut differently by the split
        # k-center
        for e, (a, b) in zip(engs, parts):
            e.cross_sims(T)
            e.cs_begin(2, None, np.arange(a, b), 10)
        nbytes = engs[0].cs_msg_bytes()
        send = [torch.zeros(nbytes, dtype=torch.uint8, device='cuda') for _ in engs]
        recv = torch.zeros(2 * nbytes, dtype=torch.uint8, device='cuda')
        for t in range(10):
            for r, e in enumerate(engs):
                e.cs_step_pack(t, send[r].data_ptr())
                e.synchronize()
            for r in range(2):
                recv[r * nbytes:(r + 1) * nbytes].copy_(send[r])
            torch.cuda.synchronize()
            for r, e in enumerate(engs):
                e.cs_step_apply_gathered(t, recv.data_ptr(), 2, r)
                e.synchronize()
        for e in engs:
            s2, v2 = e.sel_result(10)
            assert np.array_equal(s2, ref_cs)
            assert np.allclose(v2, ref_csv, rtol=1e-12, atol=1e-15)
    finally:
        for e in engs:
            e.close()


def _multi_case(seed):
    ps = (25, 25, 1)
    S, m = 3, 3
    shape = (34, 30, 4)
    allp, pools, st = [], [], np.zeros((S, 2 * m))
    rs = np.random.RandomState(seed)
    for s in range(S):
        imgs = synth_volume(shape, m, seed + 10 + s)
        allp.append(pad_imgs(imgs, ps) + [(rs.rand(*shape) > .5).astype(np.int8)])
        pools.append(list(rs.choice(int(np.prod(shape)), [110, 0, 140][s], replace=False)))
        for j in range(m):
            st[s, 2 * j], st[s, 2 * j + 1] = imgs[j].mean(), imgs[j].std()
    layers = O.pw1_layers(2)
    probe = O.normalize_batch_eval(O.get_patches(allp[0][:m], pools[0][:64], ps),
                                   [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(m)]).astype(np.float32)
    w = centered_weights(layers, (25, 25, 3), seed + 1, probe)
    return ps, S, m, allp, pools, st, layers, w, rs


def test_pw_rep_entropy_multimg(nb):
    ps, S, m, allp, pools, st, layers, w, rs = _multi_case(71)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    expr = Expr(k=8, B=40, lambda_=0., patch_shape=ps, ntb=128)
    expr.train_stats = st
    Q = nb.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'rep-entropy')
    Qo, det = O.query_rep_entropy_multimg(layers, w, allp, pools, ps, 128, st, 8, 40)
    assert len(Q) == S and len(Q[1]) == 0 and sum(len(a) for a in Q) == 8
    # the picks, as columns of the oracle's similarity matrix, must form a near-optimal greedy set: compare the
    # facility-location value of the final sets
    sel_inds = det['sel_inds']
    offs = np.concatenate([[0], np.cumsum([len(v) for v in sel_inds])])
    cols = []
    for s in range(S):
        look = {int(v): i for i, v in enumerate(sel_inds[s])}
        for v in Q[s]:
            assert int(v) in look
            cols.append(offs[s] + look[int(v)])
    val = np.sum(np.max(det['sims'][:, cols], axis=1))
    ref = np.sum(np.max(det['sims'][:, det['Q']], axis=1))
    assert abs(val / ref - 1) < 1e-3


def test_pw_core_set_multimg(nb):
    ps, S, m, allp, pools, st, layers, w, rs = _multi_case(83)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    labeled = [list(rs.choice(34 * 30 * 4, 12, replace=False)) for _ in range(S)]
    expr = Expr(k=6, B=40, lambda_=0., patch_shape=ps, ntb=128)
    expr.train_stats = st
    expr.labeled_stats = st
    Q = nb.PW_NNAL.query_multimg(expr, model, None, allp, pools, labeled, 'core-set')
    Qo, det = O.query_core_set_multimg(layers, w, allp, pools, labeled, ps, 128, st, st, 6)
    assert len(Q) == S and sum(len(a) for a in Q) == 6
    sizes = [len(p) for p in pools]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    got = np.concatenate([np.asarray(Q[s]) + offs[s] for s in range(S)]).astype(int)
    # order inside a subject follows the greedy order; compare as sets + replay tolerance on the oracle features
    assert len(np.unique(got)) == 6
    if set(got.tolist()) != set(det['Q'].tolist()):
        rep = O.kcenter_replay(det['F_u'], det['sims0'], [g for g in det['Q'] if g in got] + [g for g in got if g not in det['Q']])
        assert np.all(rep[:, 0] <= rep[:, 1] + 1e-4)
