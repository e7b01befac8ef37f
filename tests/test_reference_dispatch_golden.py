"""Golden vectors produced by the UNMODIFIED reference query dispatch (PW_NN.batch_eval, PW_NNAL.CNN_query 'entropy' and
'fi', bin_uncertainty_filter_multimg, query_multimg 'entropy') run over a fake TF session / fake cvxopt solver by
oracle/check_against_reference.py -- everything around the TF graph and the SDP solve is the reference's own code.  The
oracle must reproduce them from the stored inputs (CPU); the device path is held to them on the GPU box."""
import numpy as np
import pytest

import oracle as O
from tests.util import assert_topk_equivalent

PS, M, S = (5, 5, 1), 2, 3
LAYERS = [('conv1', [4, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']), ('fc1', [16, 'fc']), ('fc2', [12, 'fc']),
          ('fc3', [2, 'fc'])]


def _case(golden):
    w = O.he_init_weights(LAYERS, (5, 5, M), 7, bias_scale=0.1)
    imgs = golden['q_imgs']
    allp = [[imgs[s][j] for j in range(M)] + [np.zeros((12, 11, 4), dtype=np.int8)] for s in range(S)]
    pools = [list(golden['q_pool%d' % s]) for s in range(S)]
    st = golden['q_stats']
    stats0 = [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(M)]
    return w, allp, pools, st, stats0


def _case_d3(golden):
    """The d3 = 3 twin of the case: same volumes padded by one more slice in z, a 6-channel input model whose first conv is
    shrunk so that the unnormalised channels do not saturate the posteriors (oracle/check_against_reference.py)."""
    w3 = O.he_init_weights(LAYERS, (5, 5, M * 3), 8, bias_scale=0.1)
    w3['conv1'] = (w3['conv1'][0] * np.float32(0.01), w3['conv1'][1])
    imgs = golden['q_imgs']
    allp3 = [[np.pad(imgs[s][j], ((0, 0), (0, 0), (1, 1)), 'constant') for j in range(M)] + [np.zeros((12, 11, 4), dtype=np.int8)]
             for s in range(S)]
    return w3, allp3


def test_oracle_reproduces_reference_dispatch(golden):
    w, allp, pools, st, stats0 = _case(golden)
    pool0 = np.array(pools[0])
    posts = O.batch_eval(LAYERS, w, allp[0][:M], pool0, PS, 16, stats0, 'posteriors')[0]
    assert np.array_equal(posts, golden['q_posts0'])
    q, _ = O.query_entropy_single(LAYERS, w, allp[0][:M], pool0, PS, 16, stats0, 9)
    assert np.array_equal(q, golden['q_single'])
    sel, _ = O.bin_uncertainty_filter_multimg(LAYERS, w, allp, pools, PS, 16, st, 25)
    Q = O.query_entropy_multimg(LAYERS, w, allp, pools, PS, 16, st, 9)
    for s in range(S):
        assert np.array_equal(np.asarray(sel[s], dtype=np.int64), golden['q_filt%d' % s])
        assert np.array_equal(np.asarray(Q[s], dtype=np.int64), golden['q_multi%d' % s])
    qf, det = O.query_fi_sdp_single(LAYERS, w, allp[0][:M], pool0, PS, 16, stats0, 9, 30, golden['q_fi_u'], diag_load=1e-5)
    assert np.array_equal(qf, golden['q_fi_sdp_single'])
    Qf, _ = O.query_fi_sdp_multimg(LAYERS, w, allp, pools, PS, 16, st, 11, 40, golden['q_fi_u_multi'])
    for s in range(S):
        assert np.array_equal(np.asarray(Qf[s], dtype=np.int64), golden['q_fi_sdp_multi%d' % s])
    # d3 = 3: posterior pass normalised on channels 0..m-1, gradient patches block-wise (get_patches_multimg)
    w3, allp3 = _case_d3(golden)
    Q3, _ = O.query_fi_sdp_multimg(LAYERS, w3, allp3, pools, (5, 5, 3), 16, st, 11, 40, golden['q_fi_u_d3'])
    for s in range(S):
        assert np.array_equal(np.asarray(Q3[s], dtype=np.int64), golden['q_fi_sdp_multi_d3_%d' % s])
    # representativeness queries (pools without an empty subject: upstream raises on one in these branches)
    pools_r = [list(golden['q_pool_r%d' % s]) for s in range(S)]
    labeled = [list(golden['q_labeled%d' % s]) for s in range(S)]
    Qr, _ = O.query_rep_entropy_multimg(LAYERS, w, allp, pools_r, PS, 16, st, 11, 30)
    Qc, _ = O.query_core_set_multimg(LAYERS, w, allp, pools_r, labeled, PS, 16, st, st, 11)
    for s in range(S):
        assert np.array_equal(np.asarray(Qr[s], dtype=np.int64), golden['q_rep%d' % s])
        assert np.array_equal(np.asarray(Qc[s], dtype=np.int64), golden['q_cs%d' % s])
    # MC-dropout and committee queries (dropout masks of the golden run: seed 77, first pass 3, keep 0.6, FC layers)
    from oracle import mc_oracle as Mc
    for meth, key in (('MC-entropy', 'q_mc'), ('BALD', 'q_bald')):
        Qm = Mc.query_mc_multimg(LAYERS, w, allp, pools, PS, st, 11, 4, 0.6, [2, 3, 4], 77, meth, first_pass=3)[0]
        for s in range(S):
            assert np.array_equal(np.asarray(Qm[s], dtype=np.int64), golden['%s%d' % (key, s)])
    wsets = [O.he_init_weights(LAYERS, (5, 5, M), 20 + i, bias_scale=0.1) for i in range(3)]
    for meth, key in (('ensemble', 'q_ens'), ('QBC-JS', 'q_qbc')):
        Qc = Mc.query_committee_multimg(LAYERS, wsets, allp, pools, PS, 16, st, 11, meth)[0]
        for s in range(S):
            assert np.array_equal(np.asarray(Qc[s], dtype=np.int64), golden['%s%d' % (key, s)])


LAYERS_W = [('conv1', [6, 'conv', [3, 3]]), ('conv2', [5, 'conv', [5, 5]]), ('max1', [[2, 2], 'pool']),
            ('conv3', [8, 'conv', [3, 3]]), ('max2', [[2, 2], 'pool']),
            ('fc1', [40, 'fc']), ('fc2', [24, 'fc']), ('fc3', [3, 'fc'])]


def test_oracle_reproduces_reference_whole_image_dispatch(golden):
    """NNAL.CNN_query 'entropy', 'rep-entropy' and the literal multiclass 'fi' pipeline (c = 3)."""
    w = O.he_init_weights(LAYERS_W, (9, 7, 2), 6, bias_scale=0.1)
    x = golden['w_pool']
    assert np.array_equal(O.query_entropy_whole(LAYERS_W, w, x, 7)[0], golden['w_entropy'])
    assert np.array_equal(O.query_rep_entropy_whole(LAYERS_W, w, x, 7, 20)[0], golden['w_rep'])
    assert np.array_equal(O.query_fi_sdp_whole(LAYERS_W, w, x, 7, 20, golden['w_fi_u'])[0], golden['w_fi_sdp'])


@pytest.mark.gpu
def test_device_matches_reference_whole_image_dispatch(golden):
    from collections import OrderedDict
    import nnal_b200

    class Expr(object):
        pass
    w = O.he_init_weights(LAYERS_W, (9, 7, 2), 6, bias_scale=0.1)
    x = golden['w_pool']
    model = nnal_b200.NN.CNN((9, 7, 2), OrderedDict(LAYERS_W), feature_layer=len(LAYERS_W) - 2)
    model.set_weights(w)
    expr = Expr()
    expr.pars = dict(k=7, B=20, lambda_=0., batch_size=32)
    expr.pool_images = x
    post = O.forward(LAYERS_W, w, x)['posteriors']
    q = nnal_b200.NNAL.CNN_query(model, expr, np.arange(90), 'entropy', None)
    assert_topk_equivalent(q, -O.compute_entropy(post.copy()), 7, 2e-4)
    assert len(set(q.tolist()) ^ set(golden['w_entropy'].tolist())) <= 2
    q = nnal_b200.NNAL.CNN_query(model, expr, np.arange(90), 'rep-entropy', None)
    assert len(set(np.asarray(q).tolist()) ^ set(golden['w_rep'].tolist())) <= 2
    expr.pars['fi_mode'] = 'sdp'
    q, soln, sel = nnal_b200.fi.query_whole_sdp(model, expr, np.arange(90), None, return_solution=True)
    _, det = O.query_fi_sdp_whole(LAYERS_W, w, x, 7, 20, golden['w_fi_u'])
    assert soln['status'] == 'optimal' and abs(soln['primal objective'] / det['phi'] - 1) < 1e-3
    assert len(set(sel.tolist()) ^ set(det['sel'].tolist())) <= 2
    q_dev = np.array(soln['x'][:len(sel)])
    replay = sel[O.sample_query_dstr(q_dev.copy(), 7, golden['w_fi_u'])]
    assert len(set(replay.tolist()) & set(golden['w_fi_sdp'].tolist())) >= len(golden['w_fi_sdp']) - 3


@pytest.mark.gpu
def test_device_matches_reference_dispatch(golden):
    """The product's PW_NNAL.CNN_query / query_multimg on the same inputs: the reference's selections up to ties within the
    posterior tolerance; the literal 'fi' pipeline returns the reference's sample for the same uniform draws."""
    from collections import OrderedDict
    import nnal_b200

    class Expr(object):
        pass
    w, allp, pools, st, stats0 = _case(golden)
    pool0 = np.array(pools[0])
    model = nnal_b200.NN.CNN((5, 5, M), OrderedDict(LAYERS), feature_layer=len(LAYERS) - 2)
    model.set_weights(w)
    expr = Expr()
    expr.pars = dict(k=9, B=30, lambda_=0., patch_shape=PS, ntb=16, stats=stats0, img_paths=[None] * M)
    expr.train_stats, expr.nclass = st, 2
    q = nnal_b200.PW_NNAL.CNN_query(expr, model, None, allp[0][:M], pool0, None, 'entropy')
    assert_topk_equivalent(q, np.abs(golden['q_posts0'] - .5), 9, 2e-4)
    Q = nnal_b200.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'entropy')
    assert len(Q) == S and len(Q[1]) == 0
    ref = [golden['q_multi%d' % s] for s in range(S)]
    assert sum(len(a) for a in Q) == sum(len(a) for a in ref)
    assert sum(len(set(np.asarray(a).tolist()) ^ set(b.tolist())) for a, b in zip(Q, ref)) <= 2      # ties at the boundary
    expr.pars['fi_mode'] = 'sdp'
    np.random.seed(77)                                    # the generator state the golden run sampled with
    assert np.allclose(np.random.sample(9), golden['q_fi_u'])
    np.random.seed(77)
    qf, soln, sel = nnal_b200.fi.query_single_sdp(expr, model, None, allp[0][:M], pool0, return_solution=True)
    assert soln['status'] == 'optimal'
    # exact: the reference's sampler replayed on the device's q with the golden run's uniform draws
    q_dev = np.array(soln['x'][:len(sel)])
    assert np.array_equal(qf, sel[O.sample_query_dstr(q_dev.copy(), 9, golden['q_fi_u'])])
    # the SDP objective of the golden run's pipeline (float64) within the north star's 1e-3
    _, det = O.query_fi_sdp_single(LAYERS, w, allp[0][:M], pool0, PS, 16, stats0, 9, 30, golden['q_fi_u'], diag_load=1e-5)
    assert abs(soln['primal objective'] / det['phi'] - 1) < 1e-3
    # the minimiser q is not unique and the draws land in cumulative-sum bins, so single positions may differ from the
    # reference's sample; the bulk must coincide
    got, want = set(np.asarray(qf).tolist()), set(golden['q_fi_sdp_single'].tolist())
    assert got <= set(sel.tolist()) and len(got & want) >= len(want) - 3, (sorted(got), sorted(want))


@pytest.mark.gpu
def test_device_multimg_sdp_d3_block_normalisation(golden):
    """query_multimg 'fi' (literal pipeline) with a 5x5x3 patch: the device gathers the gradient patches with the
    block-wise normalisation of get_patches_multimg (every channel of modality block ch/d3), the posterior pass with
    batch_eval's channels 0..m-1 -- the A-matrices and the SDP objective of the reference's run follow."""
    from collections import OrderedDict
    import nnal_b200

    class Expr(object):
        pass
    _, _, pools, st, _ = _case(golden)
    w3, allp3 = _case_d3(golden)
    model = nnal_b200.NN.CNN((5, 5, M * 3), OrderedDict(LAYERS), feature_layer=len(LAYERS) - 2)
    model.set_weights(w3)
    expr = Expr()
    expr.pars = dict(k=11, B=40, lambda_=0., patch_shape=(5, 5, 3), ntb=16, SDP_solver='CVXOPT', fi_mode='sdp')
    expr.train_stats, expr.nclass = st, 2
    np.random.seed(79)
    Q, soln, G = nnal_b200.fi.query_multimg_sdp(expr, model, None, allp3, pools, return_solution=True)
    _, det = O.query_fi_sdp_multimg(LAYERS, w3, allp3, pools, (5, 5, 3), 16, st, 11, 40, golden['q_fi_u_d3'])
    assert soln['status'] == 'optimal'
    assert abs(soln['primal objective'] / det['phi'] - 1) < 1e-3
    # the reference's sampler replayed on the device's q with the golden run's draws: the bulk of the sample coincides
    sizes = [len(x) for x in det['sel_inds']]
    q_dev = np.array(soln['x'][:sum(sizes)])
    draws = O.sample_query_dstr(q_dev.copy(), 11, golden['q_fi_u_d3'])
    local = O.global2local_inds(draws, sizes)
    got = set()
    want = set()
    for s in range(S):
        got |= set((s, int(v)) for v in np.asarray(det['sel_inds'][s])[local[s]])
        want |= set((s, int(v)) for v in golden['q_fi_sdp_multi_d3_%d' % s])
    assert len(got & want) >= len(want) - 2, (sorted(got), sorted(want))
