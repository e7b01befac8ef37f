"""Representativeness queries: ``'rep-entropy'`` (NNAL.py:466-523, PW_NNAL.py:284-351) and ``'core-set'``
(PW_NNAL.py:353-451).  Cosine similarities come from the tcgen05
GEMM over unit-normalised feature-layer rows; the greedy loops (facility location / k-center) run on the device,
one step at a time, with one small collective per step when the pool is sharded over several ranks."""
import contextlib

import numpy as np

from . import _lib as L
from . import dist, patch_utils
from .engine import get_engine


def _stream_ctx(eng):
    if eng.message_device == 'cuda':
        import torch
        return torch.cuda.stream(torch.cuda.ExternalStream(eng.stream))
    return contextlib.nullcontext()


def _gather_rows(eng, local_pos, own, B):
    """Feature rows of the B candidate columns on every rank: each rank fills the rows it owns, a sum
    all-reduce completes the array."""
    cols = np.zeros((B, eng.feat_dim), dtype=np.float32)
    if len(local_pos):
        cols[own] = eng.pool_feature_rows(local_pos)
    if dist.is_dist():
        import torch
        t = torch.from_numpy(cols).to(dist._device())
        dist.allreduce_sum_(t)
        cols = t.cpu().numpy()
    return cols


def facility_location(eng, cols, excl_local, k):
    """k greedy steps of ``argmax_j sum_rows max(cur_row, sims[row, j])`` over this rank's rows; returns the
    selected column indices (identical on every rank) and their scores."""
    B = cols.shape[0]
    k = int(min(k, B))
    eng.rep_set(cols, excl_local, max(k, 1))
    if not dist.is_dist():
        return eng.rep_greedy(k)
    import torch
    import torch.distributed as td
    scores = torch.zeros(max(B, 1), dtype=torch.float64, device=eng.message_device)
    with _stream_ctx(eng):
        for t in range(k):
            eng.rep_step_scores(scores.data_ptr())
            td.all_reduce(scores)
            eng.rep_step_pick(t, scores.data_ptr())
    return eng.sel_result(k)


def kcenter(eng, k, gids, init):
    """k-center steps over the rows of the current pool pass; returns selected global ids and their similarities."""
    if not dist.is_dist():
        eng.cs_begin(init, None, gids, max(int(k), 1))
        sel, val = eng.cs_greedy(k)
        return sel, val
    import torch
    import torch.distributed as td
    rank, world = dist.rank_world()
    eng.cs_begin(init, None, gids, max(int(k), 1))
    nbytes = eng.cs_msg_bytes()
    send = torch.zeros(nbytes, dtype=torch.uint8, device=eng.message_device)
    recv = torch.zeros(world * nbytes, dtype=torch.uint8, device=eng.message_device)
    with _stream_ctx(eng):
        for t in range(int(k)):
            eng.cs_step_pack(t, send.data_ptr())
            td.all_gather_into_tensor(recv, send)
            eng.cs_step_apply_gathered(t, recv.data_ptr(), world, rank)
    return eng.sel_result(int(k))


def query_rep_entropy_multimg(expr, model, sess, all_padded_imgs, pool_inds):
    """``PW_NNAL.query_multimg(..., 'rep-entropy')``: per-subject local positions into ``pool_inds[s]``."""
    from .PW_NNAL import _bin_filter_core
    k, B = int(expr.pars['k']), int(expr.pars['B'])
    eng = get_engine()
    sorted_inds, _, lo, hi, sizes = _bin_filter_core(expr, model, sess, all_padded_imgs, pool_inds, B, keep=1)
    cum = np.append(-1, np.cumsum(sizes) - 1)
    set_of = cum.searchsorted(sorted_inds) - 1
    order = np.argsort(set_of, kind='stable')
    G = sorted_inds[order]                       # candidate (column) order of the reference: subject-major
    own = (G >= lo) & (G < hi)
    local = G[own] - lo
    cols = _gather_rows(eng, local, own, len(G))
    Q, _ = facility_location(eng, cols, local, min(k, len(G)))
    return patch_utils.global2local_inds(G[Q], sizes)


def query_rep_entropy_whole(model, expr, pool_inds, session):
    """``NNAL.CNN_query(..., 'rep-entropy')`` (NNAL.py:466-523): positions into ``pool_inds``."""
    from .NNAL import _posteriors_on_device
    k, B = int(expr.pars['k']), int(expr.pars['B'])
    pool_inds = np.asarray(pool_inds)
    n = len(pool_inds)
    eng, lo, hi = _posteriors_on_device(model, expr, pool_inds, session, keep=1)
    if B < n:
        eng.pool_score(L.SCORE_NEG_ENTROPY, 1e-8)
        sel_inds, _ = dist.topk_global(eng, B, lo, n)
    else:
        sel_inds = np.arange(n, dtype=np.int64)
    own = (sel_inds >= lo) & (sel_inds < hi)
    local = sel_inds[own] - lo
    cols = _gather_rows(eng, local, own, len(sel_inds))
    Q, _ = facility_location(eng, cols, local, min(k, len(sel_inds)))
    return sel_inds[Q]


def query_core_set_multimg(expr, model, sess, all_padded_imgs, pool_inds, labeled_inds):
    """``PW_NNAL.query_multimg(..., 'core-set')`` AS WRITTEN upstream: the similarity pass against the labeled
    set uses the LAST subject's labeled indices, volumes and ``expr.labeled_stats`` row only (PW_NNAL.py:387-425;
    the loop over subjects above it only rebuilds ``labeled_stats``).  Labeled volumes given as paths
    (``expr.labeled_paths != expr.train_paths``) are outside the replaced path."""
    k = int(expr.pars['k'])
    eng = get_engine()
    eng.set_model(model, sess)
    s = len(pool_inds)
    m = len(all_padded_imgs[0]) - 1
    sizes = [len(pool_inds[i]) for i in range(s)]
    n = int(np.sum(sizes))
    ps = expr.pars['patch_shape']
    # labeled features (every rank evaluates the whole labeled set: it is small)
    F_T = np.zeros((0, 0), dtype=np.float32)
    if len(labeled_inds) and len(labeled_inds[-1]):
        i = len(labeled_inds) - 1
        lst = np.array([[expr.labeled_stats[i, 2 * j], expr.labeled_stats[i, 2 * j + 1]] for j in range(m)],
                       dtype=np.float64)
        imgs = list(all_padded_imgs[i][:-1])
        eng.upload(i, imgs)
        lab = np.asarray(labeled_inds[i])
        eng.pool_begin(len(lab), 1)
        eng.pool_eval(i, lab, 0, ps, lst, L.NORM_BATCH_EVAL, shape=imgs[0].shape)
        F_T = eng.pool_feature_rows(np.arange(len(lab)))
    # pool features of this rank's block
    rank, world = dist.rank_world()
    b = dist.shard_bounds(n, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    eng.pool_begin(hi - lo, 1)
    start = 0
    for i in range(s):
        ni = sizes[i]
        a, e = max(start, lo), min(start + ni, hi)
        if e > a:
            imgs = list(all_padded_imgs[i][:-1])
            eng.upload(i, imgs)
            st = np.array([[expr.train_stats[i, 2 * j], expr.train_stats[i, 2 * j + 1]] for j in range(m)],
                          dtype=np.float64)
            eng.pool_eval(i, np.asarray(pool_inds[i])[a - start:e - start], a - lo, ps, st, L.NORM_BATCH_EVAL,
                          shape=imgs[0].shape)
        start += ni
    init = 0
    if F_T.size and hi > lo:
        eng.cross_sims(F_T)
        init = 2
    elif F_T.size:
        init = 0
    Q, _ = kcenter(eng, min(k, n), np.arange(lo, hi, dtype=np.int64), init)
    return patch_utils.global2local_inds(Q, sizes)
