"""GPU parity tests of the MC-dropout scorers (``MC-entropy``, ``BALD``; PW_NNAL.py:67-87, 232-282) against
oracle/mc_oracle.py.  The masks are a pure function of (seed, pass, layer, global pool position, unit) -- the CUDA
kernels (csrc/philox.cuh in the FC epilogue and the head kernel) and the NumPy oracle must produce the same ones, so a
single wrong mask bit shows up as an O(1) posterior error, far outside the 1e-4 tolerance."""
import os

import numpy as np
import pytest

import oracle as O
from oracle import mc_oracle as M
from tests.util import assert_topk_equivalent, centered_weights, pad_imgs, synth_volume, vol_stats

pytestmark = pytest.mark.gpu

POST_TOL = 1e-4
DROPOUT_LAYERS = [6, 7, 8]          # create_PW1, NN.py:1338


class Expr(object):
    def __init__(self, **pars):
        self.pars = pars
        self.nclass = 2


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


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


@pytest.mark.parametrize('keep', [0.5, 0.8])
def test_stochastic_batch_eval_matches_oracle(nb, keep):
    """PW_NN.batch_eval with x_feed_dict = {model.keep_prob: rate}: one stochastic pass; consecutive calls advance
    the pass counter (different masks), the same (seed, pass) reproduces the same posteriors."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(300, 70)
    model = nb.NN.create_PW1(2, dropout_rate=keep)
    model.set_weights(w)
    assert model.dropout_layers == DROPOUT_LAYERS and model.dropout_rate == keep
    eng = nb.get_engine()
    seed = 0x1234abcd5678
    eng.set_dropout_seed(seed, first_pass=11)
    feed = {model.keep_prob: model.dropout_rate}
    got = [nb.PW_NN.batch_eval(model, None, padded, pool, ps, 128, stats, 'posteriors', x_feed_dict=feed)[0] for _ in range(2)]
    pos = np.arange(len(pool))
    for t in range(2):
        want = M.batch_eval_dropout(layers, w, padded, pool, ps, stats, pos, keep, DROPOUT_LAYERS, seed, 11 + t)
        assert got[t].dtype == np.float64 and got[t].shape == want.shape
        assert np.abs(got[t] - want).max() < POST_TOL
    assert np.abs(got[0] - got[1]).max() > 1e-2                      # different passes, different masks
    eng.set_dropout_seed(seed, first_pass=11)
    again = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 64, stats, 'posteriors', x_feed_dict=feed)[0]
    assert np.array_equal(again, got[0])                             # reproducible, independent of ntb
    # no feed: deterministic posteriors, untouched by the MC machinery
    det = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 128, stats, 'posteriors')[0]
    assert np.abs(det - O.batch_eval(layers, w, padded, pool, ps, 128, stats, 'posteriors')[0]).max() < POST_TOL
    with pytest.raises(NotImplementedError):
        nb.PW_NN.batch_eval(model, None, padded, pool, ps, 128, stats, 'posteriors', x_feed_dict={'other': 1.})


def test_mc_entropy_single_volume(nb):
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(500, 80)
    keep, T, k, seed = 0.5, 5, 40, 99
    model = nb.NN.create_PW1(2, dropout_rate=keep)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_dropout_seed(seed)
    expr = Expr(k=k, B=100, lambda_=0., patch_shape=ps, ntb=128, stats=stats, MC_iters=T)
    q = nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'MC-entropy')
    av_got, ent_got = eng.pool_mc_means()
    qo, av_want = M.query_mc_single(layers, w, padded, pool, ps, stats, k, T, keep, DROPOUT_LAYERS, seed)
    assert np.abs(av_got - av_want).max() < POST_TOL
    assert q.shape == (k,)
    assert_topk_equivalent(q, np.abs(av_want - .5), k, POST_TOL)
    # the fused T-pass evaluation (conv trunk once per chunk) equals T separate stochastic batch_eval calls
    eng.set_dropout_seed(seed)
    feed = {model.keep_prob: model.dropout_rate}
    sep = [nb.PW_NN.batch_eval(model, None, padded, pool, ps, 128, stats, 'posteriors', x_feed_dict=feed)[0] for _ in range(T)]
    av_sep, ent_sep = M.mc_running_means(sep)
    assert np.abs(av_sep - av_got).max() < 1e-6 and np.abs(ent_sep - ent_got).max() < 1e-6


def test_mc_masks_independent_of_chunking_and_sharding(nb):
    """Same running means whether the pool is evaluated in one piece, in small chunks (debug option "chunk"), or as two shards
    with their global position offsets (what two ranks do)."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(333, 90)
    keep, T, seed = 0.7, 3, 7
    model = nb.NN.create_PW1(2, dropout_rate=keep)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_model(model, None)
    eng.upload(0, list(padded))
    st = np.array(stats, dtype=np.float64)

    def run(lo, hi):
        eng.set_dropout_seed(seed)
        eng.pool_mc_config(T, keep, DROPOUT_LAYERS, pos0=lo)
        try:
            eng.pool_begin(hi - lo)
            eng.pool_eval(0, pool[lo:hi], 0, ps, st, shape=padded[0].shape)
        finally:
            eng.pool_mc_config(0, 1., [])
        return eng.pool_mc_means()
    whole = run(0, len(pool))
    eng.debug_option('chunk', 100)
    try:
        chunked = run(0, len(pool))
    finally:
        eng.debug_option('chunk', 0)
    assert np.array_equal(whole[0], chunked[0]) and np.array_equal(whole[1], chunked[1])
    a, b = run(0, 150), run(150, len(pool))
    assert np.array_equal(np.concatenate([a[0], b[0]]), whole[0])
    assert np.array_equal(np.concatenate([a[1], b[1]]), whole[1])


@pytest.mark.parametrize('method', ['MC-entropy', 'BALD'])
def test_mc_queries_multimg(nb, method):
    ps = (25, 25, 1)
    S, m = 3, 3
    shape = (34, 30, 4)
    allp, pools, st = [], [], np.zeros((S, 2 * m))
    rs = np.random.RandomState(43)
    for s in range(S):
        imgs = synth_volume(shape, m, 150 + s)
        allp.append(pad_imgs(imgs, ps) + [(rs.rand(*shape) > .5).astype(np.int8)])
        pools.append(list(rs.choice(int(np.prod(shape)), [170, 0, 210][s], replace=False)))
        for j in range(m):
            st[s, 2 * j], st[s, 2 * j + 1] = imgs[j].mean(), imgs[j].std()
    layers = O.pw1_layers(2)
    probe = O.normalize_batch_eval(O.get_patches(allp[0][:m], pools[0][:64], ps),
                                   [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(m)]).astype(np.float32)
    w = centered_weights(layers, (25, 25, 3), 160, probe)
    keep, T, k, seed = 0.5, 6, 30, 2024
    model = nb.NN.create_PW1(2, dropout_rate=keep)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_dropout_seed(seed, first_pass=4)
    expr = Expr(k=k, B=100, lambda_=0., patch_shape=ps, ntb=128, MC_iters=T)
    expr.train_stats = st
    Q = nb.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, method)
    av_got, ent_got = eng.pool_mc_means()
    Qo, av_want, ent_want, scores = M.query_mc_multimg(layers, w, allp, pools, ps, st, k, T, keep, DROPOUT_LAYERS, seed, method,
                                                       first_pass=4)
    assert np.abs(av_got - av_want).max() < POST_TOL
    assert np.abs(ent_got - ent_want).max() < 1e-3
    assert len(Q) == S and len(Q[1]) == 0 and sum(len(a) for a in Q) == k
    sizes = [len(p) for p in pools]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    got_global = np.concatenate([np.asarray(Q[s]) + offs[s] for s in range(S)])
    rank_by = scores if method == 'MC-entropy' else -scores
    tol = POST_TOL if method == 'MC-entropy' else 1e-3
    kth = np.sort(rank_by)[k - 1]
    assert np.all(rank_by[got_global] <= kth + tol)
    assert np.all(np.isin(np.where(rank_by < kth - tol)[0], got_global))
    # a feed passed straight to the filter returns the concatenated posteriors of one stochastic pass (PW_NNAL.py:725-726)
    eng.set_dropout_seed(seed, first_pass=4)
    one = nb.PW_NNAL.bin_uncertainty_filter_multimg(expr, model, None, allp, pools, k, {model.keep_prob: keep})
    assert one.shape == (sum(sizes),)
    first = M.query_mc_multimg(layers, w, allp, pools, ps, st, k, 1, keep, DROPOUT_LAYERS, seed, 'MC-entropy', first_pass=4)[1]
    assert np.abs(one - first).max() < POST_TOL


@pytest.mark.parametrize('method', ['ensemble', 'QBC-JS'])
def test_committee_queries_multimg(nb, method, tmp_path):
    """'ensemble' / 'QBC-JS' in the no-label branch: committee = expr.pretrained_paths loaded into expr.model_holder
    (PW_NNAL.py:463-466), deterministic passes, same running means and scores as the MC scorers."""
    ps = (25, 25, 1)
    S, m = 2, 3
    shape = (34, 30, 4)
    allp, pools, st = [], [], np.zeros((S, 2 * m))
    rs = np.random.RandomState(47)
    for s in range(S):
        imgs = synth_volume(shape, m, 250 + s)
        allp.append(pad_imgs(imgs, ps) + [(rs.rand(*shape) > .5).astype(np.int8)])
        pools.append(list(rs.choice(int(np.prod(shape)), [190, 140][s], replace=False)))
        for j in range(m):
            st[s, 2 * j], st[s, 2 * j + 1] = imgs[j].mean(), imgs[j].std()
    layers = O.pw1_layers(2)
    probe = O.normalize_batch_eval(O.get_patches(allp[0][:m], pools[0][:64], ps),
                                   [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(m)]).astype(np.float32)
    wsets, paths = [], []
    holder = nb.NN.create_PW1(2)
    for i in range(3):
        w = centered_weights(layers, (25, 25, 3), 300 + i, probe)
        wsets.append(w)
        holder.set_weights(w)
        paths.append(str(tmp_path / ('member%d.npz' % i)))
        holder.save_weights(paths[-1])
    k = 25
    expr = Expr(k=k, B=100, lambda_=0., patch_shape=ps, ntb=128)
    expr.train_stats = st
    expr.model_holder = holder
    expr.pretrained_paths = paths
    Q = nb.PW_NNAL.query_multimg(expr, None, None, allp, pools, [[], []], method)
    av_got, ent_got = nb.get_engine().pool_mc_means()
    Qo, av_want, ent_want, scores = M.query_committee_multimg(layers, wsets, allp, pools, ps, 128, st, k, method)
    assert np.abs(av_got - av_want).max() < POST_TOL and np.abs(ent_got - ent_want).max() < 1e-3
    sizes = [len(p) for p in pools]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    got_global = np.concatenate([np.asarray(Q[s]) + offs[s] for s in range(S)])
    assert len(got_global) == k
    rank_by = scores if method == 'ensemble' else -scores
    tol = POST_TOL if method == 'ensemble' else 1e-3
    kth = np.sort(rank_by)[k - 1]
    assert np.all(rank_by[got_global] <= kth + tol)
    assert np.all(np.isin(np.where(rank_by < kth - tol)[0], got_global))
    with pytest.raises(NotImplementedError):
        nb.PW_NNAL.query_multimg(expr, None, None, allp, pools, [[1, 2], []], method)
