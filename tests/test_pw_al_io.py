"""Caller-side bookkeeping restated from PW_AL.Experiment_MultiImg.run_method (PW_AL.py:845-898): the query matrix,
the pool -> training move and the reference's file formats.  Checked against a literal transcription of the
reference loop (below) on random cases, including empty per-subject selections."""
import os

import numpy as np

import nnal_b200


def _reference_loop(Q_inds, pool_inds, training_inds):
    """PW_AL.py:856-878 transcribed verbatim (variable names kept)."""
    nQ = np.sum([len(qind) for qind in Q_inds])
    Q_mat = np.zeros((nQ, 2))
    cnt = 0
    for ind in range(len(Q_inds)):
        if len(Q_inds[ind] > 0):
            Q_mat[cnt:cnt + len(Q_inds[ind]), 0] = np.array(pool_inds[ind])[Q_inds[ind]]
            Q_mat[cnt:cnt + len(Q_inds[ind]), 1] = ind
            cnt += len(Q_inds[ind])
            training_inds[ind] += list(np.array(pool_inds[ind])[Q_inds[ind]])
            sorted_inds = -np.sort(-Q_inds[ind])
            [pool_inds[ind].pop(i) for i in sorted_inds]
    return Q_mat


def test_apply_queries_matches_reference_loop(tmp_path):
    rs = np.random.RandomState(0)
    for trial in range(5):
        S = 4
        pools = [list(rs.choice(5000, n, replace=False)) for n in (40, 0, 25, 60)]
        train = [list(rs.choice(5000, 3)) for _ in range(S)]
        Q = [rs.choice(len(pools[s]), min(len(pools[s]), k), replace=False).astype(np.int64) if len(pools[s]) else
             np.zeros(0, dtype=np.int64) for s, k in zip(range(S), (7, 0, 0, 11))]
        p1, t1 = [list(p) for p in pools], [list(t) for t in train]
        p2, t2 = [list(p) for p in pools], [list(t) for t in train]
        want = _reference_loop([q.copy() for q in Q], p1, t1)
        got = nnal_b200.PW_AL.apply_queries(Q, p2, t2)
        assert np.array_equal(got, want)
        assert p1 == p2 and t1 == t2
        nnal_b200.PW_AL.save_query_round(str(tmp_path), 'entropy', trial, got, 0.125 * trial)
        vox, subj = nnal_b200.PW_AL.load_queries(str(tmp_path), 'entropy', trial)
        assert np.array_equal(vox, want[:, 0].astype(np.int64)) and np.array_equal(subj, want[:, 1].astype(np.int64))
        # exactly what np.savetxt(q_file, Q_mat, fmt='%d') / np.savetxt(t_file, [dt]) write
        ref_q = os.path.join(str(tmp_path), 'ref_q')
        np.savetxt(ref_q, want, fmt='%d')
        assert open(ref_q).read() == open(os.path.join(str(tmp_path), 'entropy', 'queries', '%d' % trial)).read()
        assert float(open(os.path.join(str(tmp_path), 'entropy', 'AL_running_times', 'dt_%d' % trial)).read()) == 0.125 * trial


def test_weight_file_keys_follow_the_reference(tmp_path):
    """NN.save_weights writes one group per layer with datasets 'Weight' and 'Bias' (NN.py:379-396); the NPZ stand-in
    keeps those names and perform_assign_ops loads them back."""
    m = nnal_b200.NN.create_PW1(2)
    m.initialize(3)
    path = str(tmp_path / 'w.npz')
    m.save_weights(path)
    z = np.load(path)
    assert sorted(z.files) == sorted(['%s/%s' % (l, k) for l in m.weight_shapes() for k in ('Weight', 'Bias')])
    m2 = nnal_b200.NN.create_PW1(2)
    m2.perform_assign_ops(path)
    for name in m.weight_shapes():
        assert np.array_equal(m.var_dict[name][0], m2.var_dict[name][0]) and np.array_equal(m.var_dict[name][1], m2.var_dict[name][1])
