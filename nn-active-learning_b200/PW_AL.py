"""Caller-side bookkeeping of one active-learning iteration, restated from ``PW_AL.Experiment_MultiImg.run_method``
(PW_AL.py:845-898) so that the query step can be driven without TensorFlow: turn the per-subject local positions
returned by ``PW_NNAL.query_multimg`` into the reference's query matrix, move the queried voxels from the pool to the
training set, and write the ``queries/<iter>`` and ``AL_running_times/dt_<iter>`` files in the reference's formats.
Fine-tuning and evaluation stay in the reference."""
import os

import numpy as np


def apply_queries(Q_inds, pool_inds, training_inds):
    """PW_AL.py:856-878.  ``Q_inds``: list of S int arrays of local positions into ``pool_inds[s]``.  Returns ``Q_mat``
    (``[nQ, 2]``: column 0 = voxel index, column 1 = subject index) and updates ``pool_inds`` / ``training_inds``
    (lists of Python lists) in place exactly as the reference does: queried voxels are appended to the subject's
    training list and popped from its pool in descending position order."""
    nQ = int(np.sum([len(q) for q in Q_inds]))
    Q_mat = np.zeros((nQ, 2))
    cnt = 0
    for ind in range(len(Q_inds)):
        q = np.asarray(Q_inds[ind], dtype=np.int64)
        if len(q) > 0:
            vox = np.array(pool_inds[ind])[q]
            Q_mat[cnt:cnt + len(q), 0] = vox
            Q_mat[cnt:cnt + len(q), 1] = ind
            cnt += len(q)
            training_inds[ind] += list(vox)
            for i in -np.sort(-q):
                pool_inds[ind].pop(int(i))
    return Q_mat


def save_query_round(root_dir, method_name, iters, Q_mat, dt):
    """PW_AL.py:859-883: ``<root>/<method>/queries/<iters>`` (two integer columns, ``fmt='%d'``) and
    ``<root>/<method>/AL_running_times/dt_<iters>`` (one float)."""
    qdir = os.path.join(root_dir, method_name, 'queries')
    tdir = os.path.join(root_dir, method_name, 'AL_running_times')
    os.makedirs(qdir, exist_ok=True)
    os.makedirs(tdir, exist_ok=True)
    np.savetxt(os.path.join(qdir, '%d' % iters), Q_mat, fmt='%d')
    np.savetxt(os.path.join(tdir, 'dt_%d' % iters), [dt])


def load_queries(root_dir, method_name, iters):
    """Reads a ``queries/<iters>`` file back as (voxel indices, subject indices)."""
    Q = np.loadtxt(os.path.join(root_dir, method_name, 'queries', '%d' % iters), dtype=np.int64, ndmin=2)
    return Q[:, 0], Q[:, 1]
