"""Worker of tests/test_dist_cpu.py: one rank of a world_size-N gloo group running the reference-named
query functions of nnal_b200 on a NumPy fake engine.  Writes its results to <out>/rank<r>.npz."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make_case():
    from collections import OrderedDict
    import oracle as O
    from tests.util import pad_imgs, synth_volume
    ps = (5, 5, 1)
    m, S = 2, 3
    shape = (12, 11, 4)
    layers = [('conv1', [4, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']), ('fc1', [16, 'fc']), ('fc2', [12, 'fc']),
              ('fc3', [2, 'fc'])]
    w = O.he_init_weights(layers, (5, 5, m), 7, bias_scale=0.1)
    rs = np.random.RandomState(5)
    allp, pools, st = [], [], np.zeros((S, 2 * m))
    for s in range(S):
        imgs = synth_volume(shape, m, 90 + s)
        allp.append(pad_imgs(imgs, ps) + [(rs.rand(*shape) > .5).astype(np.int8)])
        pools.append(list(rs.choice(int(np.prod(shape)), [70, 0, 95][s], replace=False)))
        for j in range(m):
            st[s, 2 * j], st[s, 2 * j + 1] = imgs[j].mean(), imgs[j].std()
    return ps, m, layers, w, allp, pools, st, OrderedDict(layers)


def run(rank, world, out):
    import torch.distributed as td
    if world > 1:
        td.init_process_group('gloo', rank=rank, world_size=world)
    import nnal_b200
    from nnal_b200 import dist, engine
    from tests.fake_engine import FakeEngine
    engine._engine = FakeEngine()            # the host logic under test talks to this instead of libnnal_b200

    class Expr(object):
        pass

    ps, m, layers, w, allp, pools, st, ld = make_case()
    model = nnal_b200.NN.CNN((5, 5, m), ld, feature_layer=len(layers) - 2)
    model.set_weights(w)
    res = {}
    # single-volume queries
    expr = Expr()
    stats0 = [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(m)]
    expr.pars = dict(k=9, B=30, lambda_=0., patch_shape=ps, ntb=16, stats=stats0, fi_layers=2, fi_diag_load=1e-3)
    pool0 = np.array(pools[0])
    res['ent_single'] = nnal_b200.PW_NNAL.CNN_query(expr, model, None, allp[0][:m], pool0, None, 'entropy')
    q, obj = nnal_b200.fi.query_single(expr, model, None, allp[0][:m], pool0, return_objective=True)
    res['fi_single'], res['fi_single_obj'] = q, obj
    expr.pars['B'] = 10 ** 6                 # no pre-filter: every rank's whole block is a candidate
    q, obj = nnal_b200.fi.query_single(expr, model, None, allp[0][:m], pool0, return_objective=True)
    res['fi_all'], res['fi_all_obj'] = q, obj
    # combined round: ONE pool pass answers 'entropy' and 'fi'; and the two-pass candidate evaluation gives the same answer
    expr.pars.update(B=30)
    qe, qf = nnal_b200.PW_NNAL.CNN_query(expr, model, None, allp[0][:m], pool0, None, 'entropy+fi')
    res['combo_ent'], res['combo_fi'] = qe, qf
    engine._engine.one_pass = False
    res['fi_two_pass'] = nnal_b200.PW_NNAL.CNN_query(expr, model, None, allp[0][:m], pool0, None, 'fi')
    engine._engine.one_pass = True
    # Gram (primal) form of the objective for the selection: per-rank partial Grams, all-reduce, (d+1)^2 inverse
    expr.pars.update(B=30, fi_layers=1, fi_report=True)
    q, obj = nnal_b200.fi.query_single(expr, model, None, allp[0][:m], pool0, return_objective=True)
    rep = nnal_b200.fi.last_report
    res['fi_rep_sel'], res['fi_rep_dual_obj'] = q, np.array([obj[-1]])
    res['fi_rep'] = np.array([rep['primal_last_layer'], rep['dual_last_layer'], rep['fi_ratio'], rep['primal_reduced'],
                              rep['dual_reduced']])
    expr.pars.update(fi_layers=2, fi_report=False, B=10 ** 6)
    # the reference's literal FI pipeline (shrunk A-matrices -> SDP -> sampling): every rank evaluates the same B
    # candidates, rank 0's draw is broadcast
    expr.pars.update(B=30, fi_mode='sdp', fi_diag_load=1e-3)
    np.random.seed(1000 + rank)              # different generator states per rank: the broadcast must reconcile them
    res['fi_sdp_single'] = nnal_b200.PW_NNAL.CNN_query(expr, model, None, allp[0][:m], pool0, None, 'fi')
    # multi-volume queries
    expr = Expr()
    expr.pars = dict(k=11, B=40, lambda_=0., patch_shape=ps, ntb=16, SDP_solver='CVXOPT', fi_layers=2)
    expr.train_stats = st
    Q = nnal_b200.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'entropy')
    for s in range(len(Q)):
        res['ent_multi%d' % s] = np.asarray(Q[s])
    sel_inds, sel_posts = nnal_b200.PW_NNAL.bin_uncertainty_filter_multimg(expr, model, None, allp, pools, 25)
    for s in range(len(sel_inds)):
        res['filt_inds%d' % s] = np.asarray(sel_inds[s])
        res['filt_posts%d' % s] = np.asarray(sel_posts[s])
    Q, obj = nnal_b200.fi.query_multimg(expr, model, None, allp, pools, return_objective=True)
    for s in range(len(Q)):
        res['fi_multi%d' % s] = np.asarray(Q[s])
    res['fi_multi_obj'] = obj
    Qe, Qf = nnal_b200.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'entropy+fi')
    engine._engine.one_pass = False
    Q2 = nnal_b200.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'fi')
    engine._engine.one_pass = True
    for s in range(len(Q)):
        res['combo_multi_ent%d' % s], res['combo_multi_fi%d' % s] = np.asarray(Qe[s]), np.asarray(Qf[s])
        res['fi_multi_two_pass%d' % s] = np.asarray(Q2[s])
    expr.pars['fi_mode'] = 'sdp'
    np.random.seed(2000 + rank)
    Q = nnal_b200.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'fi')
    for s in range(len(Q)):
        res['fi_sdp_multi%d' % s] = np.asarray(Q[s])
    del expr.pars['fi_mode']
    # MC-dropout and committee queries (masks keyed by GLOBAL pool position: identical for every world size)
    mcm = nnal_b200.NN.CNN((5, 5, m), ld, feature_layer=len(layers) - 2, dropout=[[2, 3, 4], 0.6])
    mcm.set_weights(w)
    expr.pars['MC_iters'] = 4
    engine._engine.set_dropout_seed(77, first_pass=3)
    res['mc_single'] = nnal_b200.PW_NNAL.CNN_query(
        type('E', (), {'pars': dict(k=9, patch_shape=ps, ntb=16, stats=stats0, MC_iters=4)})(), mcm, None, allp[0][:m], pool0, None,
        'MC-entropy')
    for meth, key in (('MC-entropy', 'mc_multi'), ('BALD', 'bald_multi')):
        engine._engine.set_dropout_seed(77, first_pass=3)
        Q = nnal_b200.PW_NNAL.query_multimg(expr, mcm, None, allp, pools, None, meth)
        for s in range(len(Q)):
            res['%s%d' % (key, s)] = np.asarray(Q[s])
    engine._engine.set_dropout_seed(77, first_pass=3)
    res['mc_onepass'] = nnal_b200.PW_NNAL.bin_uncertainty_filter_multimg(expr, mcm, None, allp, pools, 5, {mcm.keep_prob: 0.6})
    import tempfile
    holder = nnal_b200.NN.CNN((5, 5, m), ld, feature_layer=len(layers) - 2)
    paths = []
    tmpd = tempfile.mkdtemp()
    for i in range(3):
        import oracle as O2
        holder.set_weights(O2.he_init_weights(layers, (5, 5, m), 20 + i, bias_scale=0.1))
        paths.append(os.path.join(tmpd, 'member%d_rank%d.npz' % (i, rank)))
        holder.save_weights(paths[-1])
    expr.model_holder, expr.pretrained_paths = holder, paths
    for meth, key in (('ensemble', 'ens_multi'), ('QBC-JS', 'qbc_multi')):
        Q = nnal_b200.PW_NNAL.query_multimg(expr, None, None, allp, pools, [[], [], []], meth)
        for s in range(len(Q)):
            res['%s%d' % (key, s)] = np.asarray(Q[s])
    # representativeness queries
    expr.pars['B'] = 30
    Q = nnal_b200.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'rep-entropy')
    for s in range(len(Q)):
        res['rep_multi%d' % s] = np.asarray(Q[s])
    rs2 = np.random.RandomState(17)
    labeled = [list(rs2.choice(12 * 11 * 4, 9, replace=False)) for _ in range(3)]
    expr.labeled_stats = st
    Q = nnal_b200.PW_NNAL.query_multimg(expr, model, None, allp, pools, labeled, 'core-set')
    for s in range(len(Q)):
        res['cs_multi%d' % s] = np.asarray(Q[s])
    wexpr = Expr()
    wexpr.pars = dict(k=7, B=25, lambda_=0., batch_size=32)
    wexpr.pool_images = np.random.RandomState(3).rand(90, 5, 5, m).astype(np.float32)
    res['rep_whole'] = nnal_b200.NNAL.CNN_query(model, wexpr, np.arange(90), 'rep-entropy', None)
    wexpr.pars.update(fi_mode='sdp', B=20)
    np.random.seed(3000 + rank)
    res['fi_sdp_whole'] = nnal_b200.NNAL.CNN_query(model, wexpr, np.arange(90), 'fi', None)
    # primitives
    rs = np.random.RandomState(100 + rank)
    sc = np.sort(rs.rand(6))
    pos = rs.choice(50, 6, replace=False).astype(np.int64) + 100 * rank
    res['merge_pos'], res['merge_sc'] = dist.allgather_topk(sc, pos, 8)
    res['my_sc'], res['my_pos'] = sc, pos
    res['argmin'] = np.array(dist.allreduce_argmin(float(rs.rand()), 5 + rank))
    res['bcast'] = dist.broadcast_array(np.arange(4, dtype=np.float32) + rank, world - 1)
    res['concat'] = dist.allgather_concat(np.arange(rank + 2, dtype=np.int64) + 10 * rank)
    np.savez(os.path.join(out, 'rank%d.npz' % rank), **res)
    if world > 1:
        td.barrier()
        td.destroy_process_group()


if __name__ == '__main__':
    run(int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), sys.argv[1])
