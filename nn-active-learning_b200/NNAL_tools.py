"""Drop-in for the scoring helpers of the reference's ``NNAL_tools``."""
import warnings

import numpy as np

from . import _lib as L
from .engine import get_engine


def compute_entropy(PMFs):
    """NNAL_tools.compute_entropy (NNAL_tools.py:71-85): ``PMFs`` is ``[n_classes, n_samples]``;
    zeros are bumped by 10e-8 IN PLACE in the caller's array (:80), entropy in float64."""
    PMFs[PMFs == 0] += 10e-8
    return get_engine().entropy(PMFs, L.SCORE_ENTROPY, 10e-8)


def uncertainty_filtering(posteriors, B):
    """NNAL_tools.uncertainty_filtering (NNAL_tools.py:22-36): zeros bumped by 1e-8 IN PLACE,
    indices of the B largest entropies (ties: lowest position first)."""
    posteriors[posteriors == 0] += 1e-8
    eng = get_engine()
    negH = eng.entropy(posteriors, L.SCORE_NEG_ENTROPY, 1e-8)
    return eng.topk(negH, B)


def shrink_gradient(grad, method, args=None):
    """NNAL_tools.shrink_gradient 'sum' (NNAL_tools.py:784-796) for explicit gradient lists
    (host bookkeeping; the FI path uses the factored closed form on the device)."""
    if method != 'sum':
        raise NotImplementedError("only the 'sum' mode is used by the query code")
    layer_num = int(len(grad) / 2)
    shrunk = np.zeros(layer_num)
    for t in range(layer_num):
        grW, grb = grad[2 * t], grad[2 * t + 1]
        shrunk[t] = (np.sum(grW) + np.sum(grb)) / (np.prod(grW.shape) + len(grb))
    return np.ravel(shrunk)


def sample_query_dstr(q_dstr, k, replacement=True):
    """NNAL_tools.sample_query_dstr, replacement=True branch (NNAL_tools.py:844-872)."""
    if q_dstr.min() < -.01:
        warnings.warn('Optimal q has significant negative values..')
    q_dstr[q_dstr < 0] = 0.
    if not replacement:
        raise NotImplementedError('replacement=False is unused by the query code')
    Q_inds = np.unique(q_dstr.cumsum().searchsorted(np.random.sample(k)))
    if (Q_inds == len(q_dstr)).any():
        Q_inds[Q_inds == len(q_dstr)] = len(q_dstr) - 1
    return Q_inds
