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


def SDP_query_distribution(A, lambda_, X_pool, k, tol=1e-4, max_iter=200000):
    """NNAL_tools.SDP_query_distribution (NNAL_tools.py:612-659): the query distribution minimising
    ``tr((sum_i q_i A_i)^-1)`` over the simplex; with ``lambda_ > 0`` the objective gains ``-lambda_ sum_i q_i |x_i|^2`` and
    the constraints ``X_pool q = 0`` (:625-644; ``X_pool`` [d, n] = the zero-mean refined feature matrix).  ``A`` is the list
    of tau x tau conditional FIs of ``gen_A_matrices``.  Returns a dict shaped like cvxopt's ``solvers.sdp`` solution:
    ``soln['x']`` = ``[q_1..q_n, t_1..t_tau]`` (callers slice ``soln['x'][:n]``, PW_NNAL.py:157), ``soln['status']`` =
    'optimal' when the duality certificate ``max_i tr(M^-1 A_i M^-1)/tr(M^-1) - 1`` of the returned q is at most ``2 tol`` (the
    loop stops once the certificate of the previous iterate is below ``tol``; the value reported in ``soln['gap']`` is
    recomputed for the returned q and bounds the relative distance of the objective from the SDP optimum), else
    'unknown'; plus 'primal objective', 'gap', 'iterations'."""
    A = np.asarray(A, dtype=np.float64)
    if lambda_ > 0:
        r = get_engine().sdp_query_distribution_reg(A, lambda_, np.asarray(X_pool, dtype=np.float64), tol=tol,
                                                    max_iter=min(int(max_iter), 50000))
        return _soln(r, tol)
    r = get_engine().sdp_query_distribution(A, tol=tol, max_iter=max_iter)
    return _soln(r, tol)


def _soln(r, tol):
    return {'x': np.concatenate([r['q'], r['t']]), 'status': 'optimal' if r['gap'] <= 2 * tol else 'unknown',
            'primal objective': r['objective'], 'gap': r['gap'], 'iterations': r['iterations']}


def SDP_query_distribution_from_shrunk(g, sel_posts, diag_load, k, tol=1e-4, max_iter=200000):
    """``SDP_query_distribution(gen_A_matrices(...), 0, None, k)`` without materialising the A-matrices on the host:
    ``g`` [2,B,tau] shrunk class-score gradients, ``sel_posts`` [B] = P(class 1); the A_i of PW_NNAL.py:766-814 are
    assembled on the device in the solver's layout (bit-identical to the host assembly)."""
    r = get_engine().sdp_from_shrunk(g, sel_posts, diag_load, tol=tol, max_iter=max_iter)
    return _soln(r, tol)


def solve_FIAL_SDP(A):
    """NNAL_tools.solve_FIAL_SDP (NNAL_tools.py:576-610): the same programme stated with cvxpy/MOSEK upstream
    (``expr.pars['SDP_solver'] == 'MOSEK'``, PW_NNAL.py:608-611); returns the query distribution ``q``."""
    return np.array(SDP_query_distribution(A, 0., None, None)['x'][:len(A)])


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
