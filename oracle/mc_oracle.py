"""CPU oracle (float64 NumPy) for the MC-dropout query scorers.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, line by line:
  * tf.nn.dropout on the outputs of ``model.dropout_layers`` (NN.py:167-171; PW1: layers 6, 7, 8 = fc1, fc2 and the
    logits of fc3, NN.py:1338) when ``x_feed_dict = {model.keep_prob: model.dropout_rate}`` is fed
    (PW_NNAL.py:68-69, 234-235, 248-249): kept units are divided by keep_prob, dropped units are zero
  * ``MC-entropy`` single volume (PW_NNAL.py:67-87) and multi volume (:232-244): running mean of P(class 1) over
    ``MC_iters`` stochastic passes, k smallest |mean - 0.5|
  * ``BALD`` (PW_NNAL.py:247-282): running means of P(class 1) and of the per-pass binary entropies (zeros bumped by
    1e-6), score = H(mean posterior) - mean entropy, k largest

Parity status: the arithmetic above is pure NumPy in the reference and is restated verbatim.  The MASKS cannot be
matched: the reference draws them from TensorFlow's unseeded stateful generator (no seed is set anywhere in the
repository), so no two runs of the reference agree with each other either.  This build defines the masks as a pure
function of (seed, pass, layer, global pool position, unit) through Philox4x32-10 (``philox4x32_10`` below, checked
against the published Random123 known-answer vectors in tests/test_oracle_mc.py); the CUDA path
(csrc/philox.cuh) and this file implement the same function, which is what the parity tests compare.
"""
import numpy as np

from . import nnal_oracle as O

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32 with 10 rounds (Salmon et al., SC'11).  Counter words and key words are uint32 arrays
    (broadcast against each other); returns four uint32 arrays."""
    c0, c1, c2, c3 = [np.asarray(c, dtype=np.uint32) for c in (c0, c1, c2, c3)]
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = np.uint32(k0)
    k1 = np.uint32(k1)
    with np.errstate(over='ignore'):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64)
            p1 = _M1 * c2.astype(np.uint64)
            n0 = (p1 >> np.uint64(32)).astype(np.uint32) ^ c1 ^ k0
            n1 = (p1 & _MASK32).astype(np.uint32)
            n2 = (p0 >> np.uint64(32)).astype(np.uint32) ^ c3 ^ k1
            n3 = (p0 & _MASK32).astype(np.uint32)
            c0, c1, c2, c3 = n0, n1, n2, n3
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF)
            k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return c0, c1, c2, c3


def dropout_keep_mask(seed, pass_id, site, positions, width, keep_prob):
    """Boolean keep mask ``[width, n]`` (the reference's FC activations are ``[out, N]`` column batches, NN.py:322-327)
    for dropout site ``site`` (layer index) in MC pass ``pass_id`` at the given GLOBAL pool positions:
    unit j of the sample at position s is kept iff word (j % 4) of Philox(counter = {j // 4, s, pass, site},
    key = {seed low, seed high}) < floor(keep_prob * 2^32)."""
    positions = np.asarray(positions, dtype=np.int64)
    n = len(positions)
    if keep_prob >= 1.0:
        return np.ones((width, n), dtype=bool)
    nblk = (width + 3) // 4
    blk = np.arange(nblk, dtype=np.uint32)[:, None]
    pos = (positions & 0xFFFFFFFF).astype(np.uint32)[None, :]
    r = philox4x32_10(blk, pos, np.uint32(pass_id & 0xFFFFFFFF), np.uint32(site), seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    words = np.stack(r, axis=1).reshape(nblk * 4, n)[:width]          # row 4 * blk + k  <-  word k of block blk
    thresh = np.uint64(int(keep_prob * 4294967296.0))
    return words.astype(np.uint64) < thresh


def forward_dropout(layers, weights, x, positions, keep_prob, dropout_layers, seed, pass_id, dtype=np.float64):
    """One stochastic forward pass of NN.CNN (NN.py:56-345) with tf.nn.dropout after the layers whose index is in
    ``dropout_layers`` (NN.py:167-171).  ``x`` is [N,H,W,C]; returns P(class 1) [N] and the full posteriors [c,N]."""
    r = O.forward(layers, weights, x, keep_acts=True, dtype=dtype)
    acts = r['acts']
    first = min(dropout_layers)
    if layers[first][1][1] != 'fc':
        raise ValueError('dropout sites must be FC layers (PW1: [6, 7, 8])')
    h = acts[first]['in']                       # [d, N] input of the first dropped-out layer (flattened, NN.py:296-301)
    nl = len(layers)
    for i in range(first, nl):
        name, spec = layers[i]
        W, b = weights[name]
        z = W.astype(dtype) @ h + b.astype(dtype).reshape(-1, 1)
        h = z if i == nl - 1 else np.maximum(z, 0)
        if i in dropout_layers:
            keep = dropout_keep_mask(seed, pass_id, i, positions, h.shape[0], keep_prob)
            h = np.where(keep, h / dtype(keep_prob), dtype(0))
    zmax = h.max(axis=0, keepdims=True)
    e = np.exp(h - zmax)
    post = e / e.sum(axis=0, keepdims=True)
    return post[1, :], post


def batch_eval_dropout(layers, weights, img_dat, inds, patch_shape, stats, positions, keep_prob, dropout_layers, seed,
                       pass_id):
    """PW_NN.batch_eval (PW_NN.py:357-539) for ``posteriors`` with ``x_feed_dict = {keep_prob: rate}``: gather,
    float64 normalisation, float32 feed cast, one stochastic forward.  ``positions``: global pool positions of
    ``inds`` (what the masks are keyed by)."""
    bt = O.normalize_batch_eval(O.get_patches(img_dat, np.asarray(inds), patch_shape), stats).astype(np.float32)
    return forward_dropout(layers, weights, bt, positions, keep_prob, dropout_layers, seed, pass_id)[0]


def binary_entropy_bumped(posts):
    """PW_NNAL.py:258-264: ``neg = 1 - posts`` first, then zeros of both bumped by 1e-6."""
    posts = np.array(posts, dtype=np.float64)
    neg = 1 - posts
    posts[posts == 0] += 1e-6
    neg[neg == 0] += 1e-6
    return -posts * np.log(posts) - neg * np.log(neg)


def mc_running_means(pass_posts):
    """Running means exactly as the reference writes them (PW_NNAL.py:80, 256, 266):
    ``av = (x + i * av) / (i + 1)``.  ``pass_posts``: list of T arrays of P(class 1)."""
    av_posts, av_ents = 0, 0
    for i, posts in enumerate(pass_posts):
        posts = np.asarray(posts, dtype=np.float64)
        av_posts = (posts + i * av_posts) / (i + 1)
        av_ents = (binary_entropy_bumped(posts) + i * av_ents) / (i + 1)
    return av_posts, av_ents


def mc_entropy_scores(av_posts):
    """|mean posterior - 0.5| (PW_NNAL.py:84-85, 241): k SMALLEST are queried."""
    return np.abs(av_posts - .5)


def bald_scores(av_posts, av_ents):
    """H(mean posterior) - mean entropy (PW_NNAL.py:270-278): k LARGEST are queried."""
    return binary_entropy_bumped(av_posts) - av_ents


def query_mc_single(layers, weights, padded_imgs, pool_inds, patch_shape, stats, k, T, keep_prob, dropout_layers, seed,
                    first_pass=0):
    """PW_NNAL.CNN_query 'MC-entropy' (PW_NNAL.py:67-87).  Returns (positions, av_posts)."""
    pos = np.arange(len(pool_inds))
    passes = [batch_eval_dropout(layers, weights, padded_imgs, pool_inds, patch_shape, stats, pos, keep_prob,
                                 dropout_layers, seed, first_pass + t) for t in range(T)]
    av_posts, _ = mc_running_means(passes)
    return O.stable_topk(mc_entropy_scores(av_posts), k), av_posts


def query_mc_multimg(layers, weights, all_padded_imgs, pool_inds, patch_shape, train_stats, k, T, keep_prob,
                     dropout_layers, seed, method, first_pass=0):
    """PW_NNAL.query_multimg 'MC-entropy' (:232-244) / 'BALD' (:247-282) over the concatenated multi-subject pool
    (bin_uncertainty_filter_multimg with a feed returns the concatenated posteriors, :725-726).
    Returns (per-subject local positions, av_posts, av_ents, scores)."""
    s = len(pool_inds)
    sizes = [len(pool_inds[i]) for i in range(s)]
    m = len(all_padded_imgs[0]) - 1
    starts = np.concatenate([[0], np.cumsum(sizes)])
    passes = []
    for t in range(T):
        parts = []
        for i in range(s):
            if sizes[i] == 0:
                continue
            st = [[train_stats[i, 2 * j], train_stats[i, 2 * j + 1]] for j in range(m)]
            pos = starts[i] + np.arange(sizes[i])
            parts.append(batch_eval_dropout(layers, weights, all_padded_imgs[i][:-1], pool_inds[i], patch_shape, st, pos,
                                            keep_prob, dropout_layers, seed, first_pass + t))
        passes.append(np.concatenate(parts))
    av_posts, av_ents = mc_running_means(passes)
    if method == 'MC-entropy':
        scores = mc_entropy_scores(av_posts)
        inds = O.stable_topk(scores, k)
    else:
        scores = bald_scores(av_posts, av_ents)
        inds = O.stable_topk(-scores, k)
    return O.global2local_inds(inds, sizes), av_posts, av_ents, scores


def query_committee_multimg(layers, weight_sets, all_padded_imgs, pool_inds, patch_shape, ntb, train_stats, k, method):
    """PW_NNAL.query_multimg 'ensemble' (PW_NNAL.py:453-490) / 'QBC-JS' (:492-545) in the no-label branch: the
    committee members are the weight sets of ``expr.pretrained_paths`` (:463-466), each evaluated deterministically
    (keep_prob 1, :459), posteriors and entropies averaged with the same recurrences as the MC scorers.
    Returns (per-subject local positions, av_posts, av_ents, scores)."""
    s = len(pool_inds)
    sizes = [len(pool_inds[i]) for i in range(s)]
    m = len(all_padded_imgs[0]) - 1
    passes = []
    for w in weight_sets:
        parts = []
        for i in range(s):
            if sizes[i] == 0:
                continue
            st = [[train_stats[i, 2 * j], train_stats[i, 2 * j + 1]] for j in range(m)]
            parts.append(O.batch_eval(layers, w, all_padded_imgs[i][:-1], pool_inds[i], patch_shape, ntb, st, 'posteriors')[0])
        passes.append(np.concatenate(parts))
    av_posts, av_ents = mc_running_means(passes)
    if method == 'ensemble':
        scores = mc_entropy_scores(av_posts)
        inds = O.stable_topk(scores, k)
    else:
        scores = bald_scores(av_posts, av_ents)
        inds = O.stable_topk(-scores, k)
    return O.global2local_inds(inds, sizes), av_posts, av_ents, scores
