"""Shared helpers for the parity tests."""
import numpy as np

import oracle as O


def synth_volume(shape, m, seed, dtype=np.float32):
    rs = np.random.RandomState(seed)
    return [np.clip(rs.randn(*shape) * 30 + 100, 0, None).astype(dtype) for _ in range(m)]


def pad_imgs(imgs, patch_shape):
    r = [(p - 1) // 2 for p in patch_shape]
    return [np.pad(im, ((r[0], r[0]), (r[1], r[1]), (r[2], r[2])), 'constant') for im in imgs]


def vol_stats(imgs):
    return [[float(im.mean()), float(im.std())] for im in imgs]


def centered_weights(layers, in_shape, seed, x_probe, bias_scale=0.05):
    """He-normal weights with the last-layer bias shifted so that posteriors of ``x_probe``
    straddle 0.5 (the region where selection happens)."""
    w = O.he_init_weights(layers, in_shape, seed, bias_scale=bias_scale)
    r = O.forward(layers, w, x_probe)
    z = r['output']
    name = layers[-1][0]
    W, b = w[name]
    b = b.copy()
    med = np.median(z, axis=1)
    b[:, 0] -= (med - med.mean()).astype(np.float32)
    w[name] = (W, b)
    return w


def assert_topk_equivalent(got, scores, k, tol):
    """``got`` must be a valid answer to argsort(scores)[:k] up to ties within ``tol``."""
    got = np.asarray(got)
    k = min(k, len(scores))
    assert len(got) == k and len(np.unique(got)) == k
    srt = np.sort(scores)
    kth = srt[k - 1]
    assert np.all(scores[got] <= kth + tol), 'selected a sample clearly outside the top-k'
    must = np.where(scores < kth - tol)[0]
    assert np.all(np.isin(must, got)), 'missed a sample clearly inside the top-k'
    assert np.all(np.diff(scores[got]) >= -tol), 'not in ascending score order'


def assert_feasible_for_reference_sdp(golden, q, t, tol=1e-9):
    """(q, t) against the constraint data the UNMODIFIED reference hands to cvxopt (captured by
    oracle/check_against_reference.py from NNAL_tools.SDP_query_distribution / inequality_cvx_matrix): A x = b, every
    slack h_j - G_j x positive semi-definite.  Returns the reference objective c^T x."""
    x = np.concatenate([q, t])
    assert abs((golden['sdp_Aeq'] @ x).item() - golden['sdp_beq'].item()) < 1e-9
    tau = len(t)
    for j in range(tau + 1):
        G, h = golden['sdp_G%d' % j], golden['sdp_h%d' % j]
        m = h.shape[0]
        slack = h - (G @ x).reshape(m, m)
        slack = (slack + slack.T) / 2
        assert np.linalg.eigvalsh(slack).min() > -tol * np.abs(slack).max(), 'LMI %d violated' % j
    return (golden['sdp_c'][:, 0] @ x).item()
