"""Multi-rank host logic on CPU: world_size-2 gloo groups run the reference-named query functions over a
NumPy fake engine (tests/fake_engine.py); every rank must return what a single process returns, and
that must agree with the float64 oracle."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _launch(world, out, port):
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR='127.0.0.1',
                   MASTER_PORT=str(port), OMP_NUM_THREADS='1')
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', '_dist_worker.py'), out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        o = p.communicate(timeout=600)[0].decode()
        assert p.returncode == 0, o[-3000:]
    return [np.load(os.path.join(out, 'rank%d.npz' % r)) for r in range(world)]


@pytest.fixture(scope='module')
def runs(tmp_path_factory):
    d1 = str(tmp_path_factory.mktemp('w1'))
    d2 = str(tmp_path_factory.mktemp('w2'))
    d3 = str(tmp_path_factory.mktemp('w3'))
    return _launch(1, d1, 29611)[0], _launch(2, d2, 29612), _launch(3, d3, 29613)


def test_ranks_agree_with_single_process(runs):
    one, two, three = runs
    skip = ('merge_', 'my_', 'argmin', 'bcast', 'concat')
    for multi in (two, three):
        for r in multi:
            for key in one.files:
                if key.startswith(skip):
                    continue
                a, b = one[key], r[key]
                if a.dtype.kind == 'f':
                    assert np.allclose(a, b, rtol=1e-9, atol=0), key
                else:
                    assert np.array_equal(a, b), key


def test_single_process_matches_oracle(runs):
    one = runs[0]
    sys.path.insert(0, ROOT)
    from tests._dist_worker import make_case
    ps, m, layers, w, allp, pools, st, _ = make_case()
    stats0 = [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(m)]
    pool0 = np.array(pools[0])
    q, posts = O.query_entropy_single(layers, w, allp[0][:m], pool0, ps, 16, stats0, 9)
    assert np.array_equal(one['ent_single'], q)
    qf, obj, _ = O.query_fi_single(layers, w, allp[0][:m], pool0, ps, 16, stats0, 9, 30, 2, 1e-3)
    assert np.array_equal(one['fi_single'], qf)
    assert np.allclose(one['fi_single_obj'], obj, rtol=1e-6)
    qf, obj, _ = O.query_fi_single(layers, w, allp[0][:m], pool0, ps, 16, stats0, 9, 10 ** 6, 2, 1e-3)
    assert np.array_equal(one['fi_all'], qf)
    # literal pipeline (fi_mode='sdp'): same uniform draws as the worker's rank 0 (np.random.seed(1000))
    u = np.random.RandomState(1000).random_sample(9)
    qs, det = O.query_fi_sdp_single(layers, w, allp[0][:m], pool0, ps, 16, stats0, 9, 30, u, diag_load=1e-3)
    assert np.array_equal(one['fi_sdp_single'], qs) and 0 < len(qs) <= 9
    Qs, _ = O.query_fi_sdp_multimg(layers, w, allp, pools, ps, 16, st, 11, 40, np.random.RandomState(2000).random_sample(11))
    for s in range(3):
        assert np.array_equal(one['fi_sdp_multi%d' % s], Qs[s])
    assert len(one['fi_sdp_multi1']) == 0
    Q = O.query_entropy_multimg(layers, w, allp, pools, ps, 16, st, 11)
    for s in range(3):
        assert np.array_equal(one['ent_multi%d' % s], Q[s])
    si, sp = O.bin_uncertainty_filter_multimg(layers, w, allp, pools, ps, 16, st, 25)
    for s in range(3):
        assert np.array_equal(one['filt_inds%d' % s], si[s])
        assert np.allclose(one['filt_posts%d' % s], sp[s], rtol=1e-6)      # float32 posteriors like the TF fetch
    Qf, obj, _ = O.query_fi_multimg(layers, w, allp, pools, ps, 16, st, 11, 40, 2, 1e-3)
    for s in range(3):
        assert np.array_equal(one['fi_multi%d' % s], Qf[s])
    assert np.allclose(one['fi_multi_obj'], obj, rtol=1e-6)
    from oracle import mc_oracle as M
    q, _ = M.query_mc_single(layers, w, allp[0][:m], pool0, ps, stats0, 9, 4, 0.6, [2, 3, 4], 77, first_pass=3)
    assert np.array_equal(one['mc_single'], q)
    for meth, key in (('MC-entropy', 'mc_multi'), ('BALD', 'bald_multi')):
        Qm = M.query_mc_multimg(layers, w, allp, pools, ps, st, 11, 4, 0.6, [2, 3, 4], 77, meth, first_pass=3)[0]
        for s in range(3):
            assert np.array_equal(one['%s%d' % (key, s)], Qm[s]), key
    first = M.query_mc_multimg(layers, w, allp, pools, ps, st, 11, 1, 0.6, [2, 3, 4], 77, 'MC-entropy', first_pass=3)[1]
    assert np.allclose(one['mc_onepass'], first, atol=1e-6)
    wsets = [O.he_init_weights(layers, (5, 5, m), 20 + i, bias_scale=0.1) for i in range(3)]
    for meth, key in (('ensemble', 'ens_multi'), ('QBC-JS', 'qbc_multi')):
        Qc = M.query_committee_multimg(layers, wsets, allp, pools, ps, 16, st, 11, meth)[0]
        for s in range(3):
            assert np.array_equal(one['%s%d' % (key, s)], Qc[s]), key
    Qr, _ = O.query_rep_entropy_multimg(layers, w, allp, pools, ps, 16, st, 11, 30)
    for s in range(3):
        assert np.array_equal(one['rep_multi%d' % s], Qr[s])
    labeled = [list(np.random.RandomState(17).choice(12 * 11 * 4, 9, replace=False)) for _ in range(1)]
    rs2 = np.random.RandomState(17)
    labeled = [list(rs2.choice(12 * 11 * 4, 9, replace=False)) for _ in range(3)]
    Qc, _ = O.query_core_set_multimg(layers, w, allp, pools, labeled, ps, 16, st, st, 11)
    for s in range(3):
        assert np.array_equal(one['cs_multi%d' % s], Qc[s])
    x = np.random.RandomState(3).rand(90, 5, 5, m).astype(np.float32)
    qw, _ = O.query_rep_entropy_whole(layers, w, x, 7, 25)
    assert np.array_equal(one['rep_whole'], qw)
    # whole-image literal FI pipeline (binary net here: the multiclass assembly of NNAL.py:354-414 with c = 2)
    r = O.forward(layers, w, x)
    sel = O.uncertainty_filtering(r['posteriors'].copy(), 20)
    po, go = O.shrunk_class_gradients(layers, w, x[sel])
    qo, _, _, _, _ = O.sdp_solve(O.gen_A_matrices_multiclass(po.copy(), go), 1e-4)
    draws = O.sample_query_dstr(qo.copy(), 7, np.random.RandomState(3000).random_sample(7))
    assert np.array_equal(one['fi_sdp_whole'], sel[draws])


def test_combined_round_and_two_pass_agree(runs):
    """'entropy+fi' (one pool pass) returns exactly what the two single-method queries return; evaluating the B candidates
    in a second pass (factors of the whole pool do not fit) selects the same samples as indexing them in place."""
    for r in [runs[0]] + list(runs[1]) + list(runs[2]):
        assert np.array_equal(r['combo_ent'], r['ent_single'])
        assert np.array_equal(r['combo_fi'], r['fi_single'])
        assert np.array_equal(r['fi_two_pass'], r['fi_single'])
        for s in range(3):
            assert np.array_equal(r['combo_multi_fi%d' % s], r['fi_multi%d' % s])
            assert np.array_equal(r['fi_multi_two_pass%d' % s], r['fi_multi%d' % s])
    one = runs[0]
    from tests._dist_worker import make_case
    ps, m, layers, w, allp, pools, st, _ = make_case()
    Q = O.query_entropy_multimg(layers, w, allp, pools, ps, 16, st, 11)
    for s in range(3):
        assert np.array_equal(one['combo_multi_ent%d' % s], Q[s])


def test_gram_allreduce_primal_equals_dual(runs):
    """Per-rank partial Grams summed by an all-reduce give the same primal objective tr((delta I + 2H)^-1) + (d+1)/delta
    on every rank and world size, equal to the dual (kernel) objective the greedy loop reports (last-layer FI)."""
    one, two, three = runs
    for r in [one] + list(two) + list(three):
        primal, dual, ratio, pred, dred = r['fi_rep']
        assert abs(primal / r['fi_rep_dual_obj'][0] - 1) < 1e-9
        assert abs(primal / dual - 1) < 1e-9
        assert abs(pred / dred - 1) < 1e-6            # the kernel-dependent part alone
        assert np.isfinite(ratio) and ratio > 0
        assert np.array_equal(r['fi_rep_sel'], one['fi_rep_sel'])
        assert np.allclose(r['fi_rep'], one['fi_rep'], rtol=1e-9)


def test_collective_primitives(runs):
    for multi in runs[1:]:
        world = len(multi)
        sc = np.concatenate([r['my_sc'] for r in multi])
        pos = np.concatenate([r['my_pos'] for r in multi])
        order = np.lexsort((pos, sc))[:8]
        for r in multi:
            assert np.array_equal(r['merge_pos'], pos[order]) and np.array_equal(r['merge_sc'], sc[order])
            assert np.array_equal(r['argmin'], multi[0]['argmin'])
            assert np.array_equal(r['bcast'], np.arange(4, dtype=np.float32) + world - 1)
            assert np.array_equal(r['concat'], np.concatenate([np.arange(q + 2) + 10 * q for q in range(world)]))


def test_shard_bounds():
    from nnal_b200 import dist
    for n, w in [(0, 3), (5, 8), (100, 3), (17, 1)]:
        b = dist.shard_bounds(n, w)
        assert b[0] == 0 and b[-1] == n and len(b) == w + 1
        sizes = np.diff(b)
        assert sizes.max() - sizes.min() <= 1 and np.all(sizes >= 0)
