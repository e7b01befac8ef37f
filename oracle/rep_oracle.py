"""Float64 restatement of the reference's representativeness queries (TEST ORACLE):
``rep-entropy`` (NNAL.py:466-523, PW_NNAL.py:284-351: cosine similarities between the B most uncertain samples
and the rest of the pool, greedy facility location) and ``core-set`` (PW_NNAL.py:353-451: k-center over the
feature layer), plus the similarity helpers get_self_sims / get_cross_sims (PW_NNAL.py:1041-1136)."""
import numpy as np

from .nnal_oracle import (batch_eval, bin_uncertainty_filter_multimg, global2local_inds, stable_topk,
                          uncertainty_filtering, forward)

__all__ = ['cosine_sims', 'greedy_facility_location', 'facility_location_replay', 'query_rep_entropy_whole',
           'query_rep_entropy_multimg', 'kcenter_greedy', 'kcenter_replay', 'query_core_set_multimg',
           'get_self_sims', 'get_cross_sims']


def cosine_sims(F_rem, F_unc):
    """dots / outer(norms)  (NNAL.py:493-504): ``F_rem`` [d,R], ``F_unc`` [d,B] -> [R,B]."""
    norms_rem = np.sqrt(np.sum(F_rem ** 2, axis=0))
    norms_unc = np.sqrt(np.sum(F_unc ** 2, axis=0))
    return np.dot(F_rem.T, F_unc) / np.outer(norms_rem, norms_unc)


def greedy_facility_location(sims, k):
    """The greedy loop of NNAL.py:506-521 / PW_NNAL.py:329-343 in O(k R B): at every step the candidate j
    maximising sum_rows max(sims[:, Q + [j]]) (first maximum = lowest remaining index).  Returns (Q, scores)."""
    R, B = sims.shape
    cur = np.full(R, -np.inf)
    avail = np.ones(B, dtype=bool)
    Q, sc = [], []
    for _ in range(min(k, B)):
        scores = np.sum(np.maximum(cur[:, None], sims), axis=0)
        scores[~avail] = -np.inf
        j = int(np.argmax(scores))
        Q.append(j)
        sc.append(scores[j])
        avail[j] = False
        cur = np.maximum(cur, sims[:, j])
    return np.array(Q, dtype=np.int64), np.array(sc)


def facility_location_replay(sims, Q):
    """For a given selection sequence: (score of Q[t], best available score) per step."""
    R, B = sims.shape
    cur = np.full(R, -np.inf)
    avail = np.ones(B, dtype=bool)
    out = []
    for j in Q:
        scores = np.sum(np.maximum(cur[:, None], sims), axis=0)
        scores[~avail] = -np.inf
        assert avail[j]
        out.append((scores[j], scores.max()))
        avail[j] = False
        cur = np.maximum(cur, sims[:, j])
    return np.array(out)


def query_rep_entropy_whole(layers, weights, pool_x, k, B, feature_layer=None):
    """NNAL.CNN_query 'rep-entropy' (NNAL.py:466-523) on an in-memory pool [n,H,W,C]."""
    if feature_layer is None:
        feature_layer = len(layers) - 2
    r = forward(layers, weights, pool_x.astype(np.float32), feature_layer)
    post, F = r['posteriors'], r['feature_layer']
    n = post.shape[1]
    if B < n:
        sel = uncertainty_filtering(post.copy(), B)
    else:
        B = n
        sel = np.arange(n)
    rem = np.array(sorted(set(range(n)) - set(sel.tolist())), dtype=np.int64)
    sims = cosine_sims(F[:, rem], F[:, sel])
    Q, sc = greedy_facility_location(sims, k)
    return sel[Q], {'sims': sims, 'sel': sel, 'rem': rem, 'Q': Q, 'scores': sc}


def query_rep_entropy_multimg(layers, weights, all_padded_imgs, pool_inds, patch_shape, ntb, train_stats, k, B):
    """PW_NNAL.query_multimg 'rep-entropy' (PW_NNAL.py:284-351)."""
    s = len(pool_inds)
    m = len(all_padded_imgs[0]) - 1
    F = [np.zeros((0, 0)) for _ in range(s)]
    for i in range(s):
        if len(pool_inds[i]) == 0:
            continue
        stats = [[train_stats[i, 2 * j], train_stats[i, 2 * j + 1]] for j in range(m)]
        F[i] = batch_eval(layers, weights, all_padded_imgs[i][:-1], pool_inds[i], patch_shape, ntb, stats,
                          'feature_layer')[0]
    sel_inds, _ = bin_uncertainty_filter_multimg(layers, weights, all_padded_imgs, pool_inds, patch_shape, ntb,
                                                 train_stats, B)
    F_unc = np.concatenate([F[i][:, sel_inds[i]] for i in range(s) if len(sel_inds[i]) > 0], axis=1)
    Frem = []
    for i in range(s):
        if len(pool_inds[i]) == 0:
            continue
        rem = sorted(set(range(len(pool_inds[i]))) - set(np.asarray(sel_inds[i]).tolist()))
        Frem.append(F[i][:, rem])
    F_rem = np.concatenate(Frem, axis=1)
    sims = cosine_sims(F_rem, F_unc)
    Q, sc = greedy_facility_location(sims, k)
    sizes = [len(sel_inds[i]) for i in range(s)]
    local = global2local_inds(Q, sizes)
    out = [np.array(sel_inds[i])[local[i]] for i in range(s)]
    return out, {'sims': sims, 'sel_inds': sel_inds, 'Q': Q, 'scores': sc}


def kcenter_greedy(F_u, sims0, k):
    """The k-center loop of PW_NNAL.py:437-448: q = argmin(sims) (first minimum), then
    sims = max(sims, cos(F_u[:,q], F_u)), sims[q] = inf."""
    sims = np.array(sims0, dtype=np.float64)
    norms = np.sqrt(np.sum(F_u ** 2, axis=0))
    Q, vals = [], []
    for _ in range(k):
        q = int(np.argmin(sims))
        Q.append(q)
        vals.append(sims[q])
        s_ind = np.dot(F_u[:, q].T, F_u) / (norms * norms[q])
        sims = np.maximum(sims, s_ind)
        sims[q] = np.inf
    return np.array(Q, dtype=np.int64), np.array(vals)


def kcenter_replay(F_u, sims0, Q):
    """(similarity of Q[t], smallest available similarity) per step of a given selection."""
    sims = np.array(sims0, dtype=np.float64)
    norms = np.sqrt(np.sum(F_u ** 2, axis=0))
    out = []
    for q in Q:
        out.append((sims[q], sims.min()))
        s_ind = np.dot(F_u[:, q].T, F_u) / (norms * norms[q])
        sims = np.maximum(sims, s_ind)
        sims[q] = np.inf
    return np.array(out)


def query_core_set_multimg(layers, weights, all_padded_imgs, pool_inds, labeled_inds, patch_shape, ntb,
                           train_stats, labeled_stats, k):
    """PW_NNAL.query_multimg 'core-set' (PW_NNAL.py:353-451) AS WRITTEN: the loop over labeled subjects only
    builds ``labeled_stats`` (:387-393); the similarity pass that follows it uses the LAST subject's labeled
    indices, volumes and stats only (:399-425, ``expr.labeled_paths == expr.train_paths`` branch)."""
    s = len(pool_inds)
    m = len(all_padded_imgs[0]) - 1
    Fs, sizes = [], [len(pool_inds[i]) for i in range(s)]
    for i in range(s):
        if sizes[i] == 0:
            continue
        stats = [[train_stats[i, 2 * j], train_stats[i, 2 * j + 1]] for j in range(m)]
        Fs.append(batch_eval(layers, weights, all_padded_imgs[i][:-1], pool_inds[i], patch_shape, ntb, stats,
                             'feature_layer')[0])
    F_u = np.concatenate(Fs, axis=1)
    n = F_u.shape[1]
    norms_u = np.sqrt(np.sum(F_u ** 2, axis=0))
    sims = -np.inf * np.ones(n)
    i = len(labeled_inds) - 1
    lstats = [[labeled_stats[i, 2 * j], labeled_stats[i, 2 * j + 1]] for j in range(m)]
    nT = len(labeled_inds[i])
    for b0 in range(0, nT, 1000):
        F_T = batch_eval(layers, weights, all_padded_imgs[i][:-1], np.array(labeled_inds[i])[b0:b0 + 1000],
                         patch_shape, ntb, lstats, 'feature_layer')[0]
        norms_T = np.sqrt(np.sum(F_T ** 2, axis=0))
        dots = np.dot(F_T.T, F_u)
        sims = np.max(np.concatenate((dots / np.outer(norms_T, norms_u), np.expand_dims(sims, axis=0)), axis=0),
                      axis=0)
    Q, vals = kcenter_greedy(F_u, sims, k)
    return global2local_inds(Q, sizes), {'F_u': F_u, 'sims0': sims, 'Q': Q, 'vals': vals}


def get_self_sims(F):
    """PW_NNAL.get_self_sims (PW_NNAL.py:1041-1091): max cosine similarity to any OTHER member."""
    norms = np.sqrt(np.sum(F ** 2, axis=0))
    sims = np.dot(F.T, F) / np.outer(norms, norms)
    np.fill_diagonal(sims, -np.inf)
    return np.max(sims, axis=1)


def get_cross_sims(F1, F2):
    """PW_NNAL.get_cross_sims (PW_NNAL.py:1093-1136): max cosine similarity of each F1 member to F2."""
    n1 = np.sqrt(np.sum(F1 ** 2, axis=0))
    n2 = np.sqrt(np.sum(F2 ** 2, axis=0))
    return np.max(np.dot(F1.T, F2) / np.outer(n1, n2), axis=1)
