"""Drop-in for the query dispatch of the reference's ``PW_NNAL`` (patch-wise active learning):
``CNN_query`` (PW_NNAL.py:18-166), ``query_multimg`` (:169-629) and the helpers they call.
Same method names, pool/index arguments and returned query positions."""
import numpy as np

from . import _lib as L
from . import dist, patch_utils
from .engine import get_engine


def _stats_list(stats, m):
    return np.array([[stats[j][0], stats[j][1]] for j in range(m)], dtype=np.float64)


def _score_pool_single(expr, model, sess, padded_imgs, pool_inds, keep=0, mc=None):
    """Gather + forward for (this rank's block of) a single-volume pool; leaves posteriors on
    the device.  Returns (engine, lo, hi) with [lo,hi) the block of pool positions scored here.
    ``mc = (T, keep_prob)``: T MC-dropout passes of the FC tail, running means kept on the device."""
    eng = get_engine()
    eng.set_model(model, sess)
    imgs = list(padded_imgs)
    eng.upload(0, imgs, shared=True)             # every rank is handed the same volume: PCIe carries 1/world of it per rank
    pool_inds = np.asarray(pool_inds)
    n = len(pool_inds)
    rank, world = dist.rank_world()
    b = dist.shard_bounds(n, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    if mc is not None:
        eng.pool_mc_config(mc[0], mc[1], model.dropout_layers, pos0=lo)
    try:
        eng.pool_begin(hi - lo, keep)
        eng.pool_eval(0, pool_inds[lo:hi], 0, expr.pars['patch_shape'],
                      _stats_list(expr.pars['stats'], len(imgs)), L.NORM_BATCH_EVAL, shape=imgs[0].shape)
    finally:
        if mc is not None:
            eng.pool_mc_config(0, 1., [])
    return eng, lo, hi


def binary_uncertainty_filter(posts, B):
    """PW_NNAL.binary_uncertainty_filter (PW_NNAL.py:671-681)."""
    eng = get_engine()
    return eng.topk(np.abs(np.array(posts, dtype=np.float64) - 0.5), B)


def CNN_query(expr, model, sess, padded_imgs, pool_inds, tr_inds, method_name):
    """PW_NNAL.CNN_query (PW_NNAL.py:18-166).  Returns positions into ``pool_inds``.

    ``entropy``: k pool samples with the smallest |P(class 1) - 0.5| (:51-65), ascending.
    ``fi``: uncertainty pre-filter to B (:98-115), conditional FI of the last FC layers in
    factored form, deterministic greedy selection of k instead of the
    reference's SDP + random sampling (:146-163); returns ``sel_inds[Q]``.
    ``expr.pars['fi_mode'] = 'sdp'`` runs the reference's own pipeline instead: shrunk-coordinate
    A-matrices (:133-137), SDP query distribution (:154-157), ``sample_query_dstr`` (:160-163)."""
    if method_name == 'random':
        n = len(pool_inds)
        return np.random.permutation(n)[:expr.pars['k']]

    if method_name == 'entropy':
        k = expr.pars['k']
        eng, lo, hi = _score_pool_single(expr, model, sess, padded_imgs, pool_inds)
        eng.pool_score(L.SCORE_BINARY)
        q, _ = dist.topk_global(eng, k, lo, len(pool_inds))
        return q

    if method_name == 'MC-entropy':
        # PW_NNAL.py:67-87: running mean of P(class 1) over MC_iters dropout passes, k smallest |mean - 0.5|
        k = expr.pars['k']
        eng, lo, hi = _score_pool_single(expr, model, sess, padded_imgs, pool_inds,
                                         mc=(int(expr.pars['MC_iters']), float(model.dropout_rate)))
        eng.pool_score(L.SCORE_MC_BINARY)
        q, _ = dist.topk_global(eng, k, lo, len(pool_inds))
        return q

    if method_name == 'fi':
        from . import fi
        if expr.pars.get('fi_mode', 'greedy') == 'sdp':
            return fi.query_single_sdp(expr, model, sess, padded_imgs, pool_inds)
        return fi.query_single(expr, model, sess, padded_imgs, pool_inds)

    if method_name == 'entropy+fi':
        # not a reference method name: ONE pool pass answering both queries (a query round as SURVEY.md 8d defines it);
        # returns (q_entropy, q_fi), each exactly what the single-method call returns
        from . import fi
        return fi.query_single(expr, model, sess, padded_imgs, pool_inds, also_entropy=True)

    raise NotImplementedError('query method %r is not part of the replaced path' % method_name)


def _A_from_shrunk(g, sel_posts, diag_load, as_list=True):
    """The A-matrix assembly of PW_NNAL.gen_A_matrices (PW_NNAL.py:766-814) from shrunk gradients ``g`` [2,B,tau]:
    p < 1e-6 -> p = 0 and only g0 (:770-780), p > 1-1e-6 -> p = 1 and only g1 (:782-793),
    ``A_i = (1-p) g0 g0^T + p g1 g1^T + diag_load I`` (:810-814).  Vectorised over the B samples with the
    reference's order of operations (bit-identical to the per-sample loop).  ``as_list=False`` keeps the [B,tau,tau]
    array (the SDP entry point takes either; 10,000 array views cost more than the arithmetic)."""
    tau = g.shape[2]
    p = np.array(sel_posts, dtype=np.float64)
    lo, hi = p < 1e-6, p > 1 - 1e-6
    p[lo], p[hi] = 0., 1.
    g0 = np.where(hi[:, None], 0., g[0])
    g1 = np.where(lo[:, None], 0., g[1])
    B = len(p)
    A = np.einsum('bi,bj->bij', g0, g0).reshape(B, tau * tau)
    A1 = np.einsum('bi,bj->bij', g1, g1).reshape(B, tau * tau)
    A *= (1. - p)[:, None]
    A1 *= p[:, None]
    A += A1
    A += (np.eye(tau) * diag_load).reshape(1, tau * tau)
    A = A.reshape(B, tau, tau)
    return list(A) if as_list else A


def gen_A_matrices(expr, model, sess, sel_patches, sel_posts, diag_load=1e-5):
    """PW_NNAL.gen_A_matrices (PW_NNAL.py:738-816): conditional FIs of the (already normalised) ``sel_patches``
    [B,d1,d2,m*d3] in the reference's shrunk coordinates, ``A_i = (1-p) g0 g0^T + p g1 g1^T + diag_load I`` with
    ``g_y = shrink_gradient(d log p_y / d theta, 'sum')`` over all trainable layers.  The 2B single-sample
    ``sess.run(model.grad_posts[y])`` calls are one batched backward pass on the device (csrc/shrunk.cu)."""
    eng = get_engine()
    eng.set_model(model, sess)
    if eng.n_class != 2:
        raise NotImplementedError('gen_A_matrices assumes binary classification (PW_NNAL.py:765)')
    _, g = eng.fi_shrunk_images(np.asarray(sel_patches))
    return _A_from_shrunk(g, np.asarray(sel_posts, dtype=np.float64), diag_load)


def _pool_pass_multimg(expr, model, sess, all_padded_imgs, pool_inds, keep=0, mc=None):
    """Pool pass over (this rank's block of) the concatenated multi-subject pool.  ``mc = (T, keep_prob)`` runs T
    MC-dropout passes of the FC tail per chunk.  Returns (engine, lo, hi, per-subject sizes, n)."""
    eng = get_engine()
    eng.set_model(model, sess)
    s = len(pool_inds)
    img_ind_sizes = [len(pool_inds[i]) for i in range(s)]
    m = len(all_padded_imgs[0]) - 1
    n = int(np.sum(img_ind_sizes))
    rank, world = dist.rank_world()
    b = dist.shard_bounds(n, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    if mc is not None:
        eng.pool_mc_config(mc[0], mc[1], model.dropout_layers, pos0=lo)
    try:
        eng.pool_begin(hi - lo, keep)
        start = 0
        for i in range(s):
            ni = img_ind_sizes[i]
            a, e = max(start, lo), min(start + ni, hi)
            if e > a:
                imgs = list(all_padded_imgs[i][:-1])
                eng.upload(i, imgs)
                stats = np.array([[expr.train_stats[i, 2 * j], expr.train_stats[i, 2 * j + 1]] for j in range(m)],
                                 dtype=np.float64)
                inds_i = np.asarray(pool_inds[i])[a - start:e - start]
                eng.pool_eval(i, inds_i, a - lo, expr.pars['patch_shape'], stats, L.NORM_BATCH_EVAL,
                              shape=imgs[0].shape)
            start += ni
    finally:
        if mc is not None:
            eng.pool_mc_config(0, 1., [])
    return eng, lo, hi, img_ind_sizes, n


def _bin_filter_core(expr, model, sess, all_padded_imgs, pool_inds, B, keep=0):
    """Pool pass over (this rank's block of) the concatenated multi-subject pool + global top-B by
    |p - 0.5|.  Returns (sorted global positions, their posteriors, lo, hi, per-subject sizes)."""
    eng, lo, hi, img_ind_sizes, n = _pool_pass_multimg(expr, model, sess, all_padded_imgs, pool_inds, keep)
    eng.pool_score(L.SCORE_BINARY)
    post_local = eng.pool_posteriors()[1, :].astype(np.float64)
    sorted_inds, _ = dist.topk_global(eng, B, lo, n)
    # posteriors of the selected samples (owners contribute theirs)
    mine = (sorted_inds >= lo) & (sorted_inds < hi)
    sel_p = np.zeros(len(sorted_inds))
    sel_p[mine] = post_local[sorted_inds[mine] - lo]
    if dist.is_dist():
        import torch
        t = torch.from_numpy(sel_p).to(dist._device())
        dist.allreduce_sum_(t)
        sel_p = t.cpu().numpy()
    return sorted_inds, sel_p, lo, hi, img_ind_sizes


def bin_uncertainty_filter_multimg(expr, model, sess, all_padded_imgs, pool_inds, B, x_feed_dict={}):
    """PW_NNAL.bin_uncertainty_filter_multimg (PW_NNAL.py:684-736): posteriors of every
    subject's pool, rank |p - 0.5| over the CONCATENATED pool, keep B, split back with
    global2local_inds.  Returns ``(sel_inds, sel_posts)`` lists per subject."""
    if len(x_feed_dict) > 0:
        # PW_NNAL.py:725-726: with a feed the reference returns the raw concatenated posteriors of this
        # (stochastic) pass instead of a selection
        from .PW_NN import keep_prob_from_feed
        keep_prob = keep_prob_from_feed(model, x_feed_dict)
        eng, lo, hi, _, n = _pool_pass_multimg(expr, model, sess, all_padded_imgs, pool_inds, mc=(1, keep_prob))
        return dist.allgather_concat(eng.pool_posteriors()[1, :].astype(np.float64))
    sorted_inds, sel_p, _, _, img_ind_sizes = _bin_filter_core(expr, model, sess, all_padded_imgs, pool_inds, B)
    s = len(pool_inds)
    sel_inds = patch_utils.global2local_inds(sorted_inds, img_ind_sizes)
    cum = np.append(-1, np.cumsum(img_ind_sizes) - 1)
    set_of = cum.searchsorted(sorted_inds) - 1
    sel_posts = [sel_p[set_of == i] for i in range(s)]
    return sel_inds, sel_posts


def query_multimg(expr, model, sess, all_padded_imgs, pool_inds, labeled_inds, method_name):
    """PW_NNAL.query_multimg (PW_NNAL.py:169-629): list of S arrays of local positions into
    ``pool_inds[s]`` (possibly empty)."""
    k = expr.pars['k']
    img_ind_sizes = [len(pool_inds[i]) for i in range(len(pool_inds))]

    if method_name == 'random':
        npool = np.sum(img_ind_sizes)
        inds = np.random.permutation(npool)[:k]
        return patch_utils.global2local_inds(inds, img_ind_sizes)

    if method_name == 'entropy':
        return bin_uncertainty_filter_multimg(expr, model, sess, all_padded_imgs, pool_inds, k)[0]

    if method_name in ('MC-entropy', 'BALD'):
        # PW_NNAL.py:232-244 (MC-entropy): k smallest |mean_t P_t - 0.5|;  :247-282 (BALD): k largest
        # H(mean_t P_t) - mean_t H(P_t).  The T passes share one evaluation of the conv trunk per chunk.
        eng, lo, hi, _, n = _pool_pass_multimg(expr, model, sess, all_padded_imgs, pool_inds,
                                               mc=(int(expr.pars['MC_iters']), float(model.dropout_rate)))
        eng.pool_score(L.SCORE_MC_BINARY if method_name == 'MC-entropy' else L.SCORE_NEG_BALD)
        inds, _ = dist.topk_global(eng, k, lo, n)
        return patch_utils.global2local_inds(inds, img_ind_sizes)

    if method_name in ('ensemble', 'QBC-JS'):
        # PW_NNAL.py:453-490 (ensemble): k smallest |mean_i P_i - 0.5| over the committee;  :492-545 (QBC-JS): k largest
        # H(mean_i P_i) - mean_i H(P_i).  With no labels yet the committee is expr.pretrained_paths loaded into
        # expr.model_holder (:463-466); with labels the reference fine-tunes the previous model once per member
        # (:467-476) -- training, which stays in the reference.
        n_labels = np.sum([len(labeled_inds[i]) for i in range(len(labeled_inds))]) if labeled_inds is not None else 0
        if n_labels > 0:
            raise NotImplementedError('committee members are fine-tuned by the reference (PW_AL.finetune_multimg); '
                                      'load their weights into expr.pretrained_paths to score them here')
        eng = None
        try:
            for i in range(len(expr.pretrained_paths)):
                expr.model_holder.perform_assign_ops(expr.pretrained_paths[i], sess)
                eng, lo, hi, _, n = _pool_pass_multimg(expr, expr.model_holder, sess, all_padded_imgs, pool_inds)
                eng.pool_ensemble_accumulate(i)
        finally:
            if eng is not None:
                eng.pool_ensemble_end()
        eng.pool_score(L.SCORE_MC_BINARY if method_name == 'ensemble' else L.SCORE_NEG_BALD)
        inds, _ = dist.topk_global(eng, k, lo, n)
        return patch_utils.global2local_inds(inds, img_ind_sizes)

    if method_name == 'fi':
        from . import fi
        if expr.pars.get('fi_mode', 'greedy') == 'sdp':
            return fi.query_multimg_sdp(expr, model, sess, all_padded_imgs, pool_inds)
        return fi.query_multimg(expr, model, sess, all_padded_imgs, pool_inds)

    if method_name == 'entropy+fi':
        from . import fi
        return fi.query_multimg(expr, model, sess, all_padded_imgs, pool_inds, also_entropy=True)

    if method_name == 'rep-entropy':
        from . import rep
        return rep.query_rep_entropy_multimg(expr, model, sess, all_padded_imgs, pool_inds)

    if method_name == 'core-set':
        from . import rep
        return rep.query_core_set_multimg(expr, model, sess, all_padded_imgs, pool_inds, labeled_inds)

    raise NotImplementedError('query method %r is not part of the replaced path' % method_name)


def stoch_approx_IF(model, sess, tr_patches, pool_patches, max_iter, scale=50):
    """PW_NNAL.stoch_approx_IF (PW_NNAL.py:851-881): stochastic approximation of the pool samples' influence through the
    last FC layer.  ``grads`` = NN.LLFC_grads of the pool at its predicted (weak) labels; ``V_0 = grads``; per iteration a
    random training patch ``np.random.randint(ntr)`` (NumPy's global generator, one draw per iteration as upstream, :873) and
    ``V <- grads + V - H V / scale`` with ``H = -NN.LLFC_hess`` of it.  Returns ``(V_t [(d+1)c, n], weak_labels)``.

    The pool and the drawn training patches go through ONE batched forward pass each on the device, and the recursion runs
    there in factored form (``nnal_if_lissa``): the reference builds a ((d+1)c)^2 Hessian per iteration."""
    from .NN import _last_layer_factors
    from .engine import get_engine
    tr_patches = np.asarray(tr_patches)
    ntr = tr_patches.shape[0]
    pool_post, pool_U = _last_layer_factors(model, sess, {model.x: pool_patches})
    weak_labels = np.argmax(pool_post, axis=0)
    draws = np.array([np.random.randint(ntr) for _ in range(int(max_iter))], dtype=np.int64)
    uniq, inv = np.unique(draws, return_inverse=True)
    if len(uniq):
        tp, tU = _last_layer_factors(model, sess, {model.x: tr_patches[uniq]})
        tr_post, tr_U = tp.T[inv], tU.T[inv]
    else:
        tr_post = np.zeros((0, pool_post.shape[0]), dtype=np.float32)
        tr_U = np.zeros((0, pool_U.shape[0]), dtype=np.float32)
    V = get_engine().if_lissa(pool_post, pool_U.T, weak_labels, tr_post, tr_U, float(scale))
    return V, weak_labels


def refine_feature_matrix(F, B):
    """PW_NNAL.refine_feature_matrix (PW_NNAL.py:819-849): rows of the feature matrix ``F`` [d, n] that make it full row-rank
    with a moderate condition number -- the int(B/2) rows with the most positive entries, then the tail of that list is
    dropped until the rank is full and until cond <= 1e6 (never below one row).  Host NumPy like upstream (the matrix is
    B/2 x B at most); returns a copy."""
    order = np.argsort(-np.sum(F > 0, axis=1))[:int(B / 2)]
    while np.linalg.matrix_rank(F[order, :]) < len(order):
        order = order[:-1]
    while np.linalg.cond(F[order, :]) > 1e6:
        order = order[:-1]
        if len(order) == 1:
            break
    return F[order, :].copy()
