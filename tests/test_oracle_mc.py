"""CPU tests of the MC-dropout oracle (oracle/mc_oracle.py): the Philox4x32-10 generator against the published
Random123 known-answer vectors, the mask definition, and the reference's running-mean / BALD arithmetic
(PW_NNAL.py:67-87, 232-282)."""
import numpy as np

import oracle as O
from oracle import mc_oracle as M


def _philox_scalar(ctr, key):
    """Independent pure-Python restatement (arbitrary-precision ints) of Philox4x32-10."""
    c = list(ctr)
    k = list(key)
    for _ in range(10):
        p0 = 0xD2511F53 * c[0]
        p1 = 0xCD9E8D57 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k[1], p0 & 0xFFFFFFFF]
        k = [(k[0] + 0x9E3779B9) & 0xFFFFFFFF, (k[1] + 0xBB67AE85) & 0xFFFFFFFF]
    return c


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 10 rounds."""
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kats:
        assert tuple(_philox_scalar(ctr, key)) == want
        got = M.philox4x32_10(*[np.array([c], dtype=np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_philox_vectorised_matches_scalar():
    rs = np.random.RandomState(0)
    c = rs.randint(0, 2 ** 32, size=(4, 50), dtype=np.uint64).astype(np.uint32)
    k0, k1 = 0x12345678, 0x9abcdef0
    got = M.philox4x32_10(c[0], c[1], c[2], c[3], k0, k1)
    for i in range(50):
        want = _philox_scalar([int(c[j, i]) for j in range(4)], [k0, k1])
        assert [int(g[i]) for g in got] == want


def test_keep_mask_definition_and_invariance():
    seed, keep = 0xdeadbeefcafe, 0.7
    pos = np.arange(1000, 1400)
    m = M.dropout_keep_mask(seed, 3, 7, pos, 4096, keep)
    assert m.shape == (4096, 400) and abs(m.mean() - keep) < 5e-3
    # a pure function of (seed, pass, site, position, unit): any subset / order of positions gives the same columns
    sub = np.array([1399, 1000, 1234])
    assert np.array_equal(M.dropout_keep_mask(seed, 3, 7, sub, 4096, keep), m[:, sub - 1000])
    # unit j uses word j % 4 of block j // 4; widths that are not a multiple of 4 truncate
    assert np.array_equal(M.dropout_keep_mask(seed, 3, 7, pos, 2, keep), m[:2])
    # different pass / site / seed -> different masks
    for other in (M.dropout_keep_mask(seed, 4, 7, pos, 4096, keep), M.dropout_keep_mask(seed, 3, 6, pos, 4096, keep),
                  M.dropout_keep_mask(seed + 1, 3, 7, pos, 4096, keep)):
        assert abs((other == m).mean() - (keep ** 2 + (1 - keep) ** 2)) < 5e-3
    assert M.dropout_keep_mask(seed, 0, 0, pos, 8, 1.0).all()
    # threshold is floor(keep * 2^32) on the raw word
    w = M.philox4x32_10(np.uint32(0), np.uint32(1000), np.uint32(3), np.uint32(7), seed & 0xffffffff, seed >> 32)
    assert bool(m[1, 0]) == (int(w[1]) < int(keep * 2 ** 32))


def test_forward_dropout_semantics():
    layers = [('conv1', [4, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']), ('fc1', [16, 'fc']), ('fc2', [8, 'fc']),
              ('fc3', [2, 'fc'])]
    w = O.he_init_weights(layers, (6, 6, 2), 0, bias_scale=0.1)
    x = np.random.RandomState(1).randn(5, 6, 6, 2)
    pos = np.arange(5) + 40
    # keep_prob = 1: identity
    p1, post = M.forward_dropout(layers, w, x, pos, 1.0, [2, 3, 4], 9, 0)
    assert np.allclose(post, O.forward(layers, w, x)['posteriors'], atol=1e-15)
    # explicit restatement with the masks
    keep = 0.6
    r = O.forward(layers, w, x, keep_acts=True)
    h = r['acts'][2]['in']
    for i in (2, 3, 4):
        W, b = w[layers[i][0]]
        z = W.astype(np.float64) @ h + b.astype(np.float64).reshape(-1, 1)
        h = z if i == 4 else np.maximum(z, 0)
        h = h * M.dropout_keep_mask(9, 5, i, pos, h.shape[0], keep) / keep
    e = np.exp(h - h.max(0))
    want = (e / e.sum(0))[1]
    got = M.forward_dropout(layers, w, x, pos, keep, [2, 3, 4], 9, 5)[0]
    assert np.allclose(got, want, atol=1e-14)
    assert not np.allclose(got, M.forward_dropout(layers, w, x, pos, keep, [2, 3, 4], 9, 6)[0])


def test_running_means_and_bald():
    rs = np.random.RandomState(2)
    passes = [rs.rand(50) for _ in range(7)]
    passes[3][:4] = [0., 1., 0., 1.]                    # exercise the 1e-6 zero bumps
    av_p, av_e = M.mc_running_means(passes)
    assert np.allclose(av_p, np.mean(passes, axis=0), atol=1e-15)

    def ent(p):
        p = np.array(p)
        q = 1 - p
        p[p == 0] += 1e-6
        q[q == 0] += 1e-6
        return -p * np.log(p) - q * np.log(q)
    assert np.allclose(av_e, np.mean([ent(p) for p in passes], axis=0), atol=1e-15)
    s = M.bald_scores(av_p, av_e)
    assert np.all(s > -1e-12)                            # Jensen: H(mean) >= mean H
    assert np.allclose(M.mc_entropy_scores(av_p), np.abs(av_p - .5))
    # the caller's arrays are not modified (the reference bumps copies created inside the loop)
    assert passes[3][0] == 0.
