"""GPU parity tests (run on the B200 box: ``pytest -m gpu``).  Every test goes through the
reference-named shims of ``nnal_b200`` and hence through the C ABI of libnnal_b200.so, and
compares against the float64 oracle / the golden vectors produced by the reference's own
NumPy helpers."""
import numpy as np
import pytest

import oracle as O
from tests.util import (assert_topk_equivalent, centered_weights, pad_imgs, synth_volume,
                        vol_stats)

pytestmark = pytest.mark.gpu

POST_TOL = 1e-4      # north_star: posteriors within 1e-4 absolute


class Expr(object):
    def __init__(self, **pars):
        self.pars = pars
        self.nclass = 2


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


# ------------------------------------------------------------------ gather
def test_gather_golden_bitexact(nb, golden):
    for ci in range(4):
        imgs = list(golden['gather%d_imgs' % ci])
        out = nb.patch_utils.get_patches(imgs, golden['gather%d_inds' % ci], tuple(golden['gather%d_pshape' % ci]))
        assert out.dtype == np.float64 and out.shape == golden['gather%d_out' % ci].shape
        assert np.array_equal(out, golden['gather%d_out' % ci])


@pytest.mark.parametrize('dtype', [np.float32, np.float64, np.int16, np.uint8])
@pytest.mark.parametrize('ps', [(25, 25, 1), (5, 3, 3), (1, 1, 1)])
def test_gather_vs_oracle(nb, dtype, ps):
    rs = np.random.RandomState(7)
    shape = (40, 37, 9)
    imgs = [(rs.rand(*shape) * 200).astype(dtype) for _ in range(3)]
    padded = pad_imgs(imgs, ps)
    mask = (rs.rand(*shape) > .5).astype(np.int8)
    inds = rs.choice(int(np.prod(shape)), 301, replace=False)
    inds[:3] = [0, np.prod(shape) - 1, shape[2] - 1]
    ref = O.get_patches(padded, inds, ps)
    assert np.array_equal(nb.patch_utils.get_patches(padded, inds, ps), ref)
    got, lab = nb.patch_utils.get_patches(imgs, inds, ps, False, mask)
    rref, rlab = O.get_patches(imgs, inds, ps, False, mask)
    assert np.array_equal(got, rref) and np.array_equal(lab, rlab)


def test_gather_edge_cases(nb):
    imgs = synth_volume((8, 8, 4), 2, 1)
    padded = pad_imgs(imgs, (3, 3, 1))
    out = nb.patch_utils.get_patches(padded, np.array([], dtype=np.int64), (3, 3, 1))
    assert out.shape == (0, 3, 3, 2)
    with pytest.raises(ValueError):
        nb.patch_utils.get_patches(padded, np.array([8 * 8 * 4]), (3, 3, 1))     # unravel_index raises
    one = nb.patch_utils.get_patches(padded, np.array([5]), (3, 3, 1))
    assert np.array_equal(one, O.get_patches(padded, np.array([5]), (3, 3, 1)))


def test_gather_multimg_golden(nb, golden):
    imgs, masks = golden['multi_imgs'], golden['multi_masks']
    allp = [list(imgs[s]) + [masks[s]] for s in range(imgs.shape[0])]
    inds = [list(golden['multi_inds%d' % s]) for s in range(3)]
    p, l = nb.patch_utils.get_patches_multimg(allp, inds, (5, 5, 1), golden['multi_stats'])
    for s in range(3):
        if len(inds[s]) == 0:
            assert len(p[s]) == 0
            continue
        assert np.array_equal(p[s], golden['multi_out%d' % s])     # float64 normalisation, bit-exact
        assert np.array_equal(l[s], golden['multi_labels%d' % s])


# ------------------------------------------------------------------ entropy / top-k
def test_entropy_golden(nb, golden):
    P = golden['entropy_P'].copy()
    H = nb.NNAL_tools.compute_entropy(P)
    assert P[1, 3] == 10e-8                                         # in-place side effect kept
    assert np.allclose(H, golden['entropy_H'], rtol=1e-12, atol=1e-15)
    P = golden['entropy_P'].copy()
    sel = nb.NNAL_tools.uncertainty_filtering(P, 9)
    assert np.array_equal(sel, golden['unc_sel'])
    assert np.array_equal(nb.PW_NNAL.binary_uncertainty_filter(golden['bin_posts'], 20), golden['bin_sel'])


@pytest.mark.parametrize('n,k', [(1, 1), (5, 10), (1000, 100), (70000, 10000), (1 << 20, 100), (300000, 20000)])
def test_topk_exact_with_ties(nb, n, k):
    rs = np.random.RandomState(n % 97)
    s = rs.rand(n)
    s[rs.rand(n) < .3] = 0.5          # massive ties (saturated posteriors give |p-.5| = .5)
    s[rs.rand(n) < .05] = 0.0
    eng = nb.get_engine()
    got = eng.topk(s, k)
    assert np.array_equal(got, np.argsort(s, kind='stable')[:k])


def test_pixelwise_entropy_config4_shape(nb):
    """config 4(ii) at reduced z-extent: [c=2,256,256,z] posteriors -> entropy map."""
    rs = np.random.RandomState(5)
    z = rs.randn(2, 256, 256, 6)
    p = np.exp(z) / np.exp(z).sum(0, keepdims=True)
    p = p.astype(np.float32).astype(np.float64)
    H = nb.NNAL_tools.compute_entropy(p.reshape(2, -1).copy()).reshape(256, 256, 6)
    assert np.allclose(H, O.pixelwise_entropy(p), rtol=1e-12)


# ------------------------------------------------------------------ forward
SMALL = [('conv1', [6, 'conv', [3, 3]]), ('conv2', [5, 'conv', [5, 5]]), ('max1', [[2, 2], 'pool']),
         ('conv3', [8, 'conv', [3, 3]]), ('max2', [[2, 2], 'pool']),
         ('fc1', [40, 'fc']), ('fc2', [24, 'fc']), ('fc3', [3, 'fc'])]


def _model(nb, layers, in_shape, w, feature_layer=None):
    from collections import OrderedDict
    m = nb.NN.CNN(in_shape, OrderedDict(layers), feature_layer=feature_layer)
    m.set_weights(w)
    return m


def test_small_net_whole_image_entropy(nb):
    """NNAL.CNN_query 'entropy' on a generic layer dictionary (odd sizes, c=3)."""
    rs = np.random.RandomState(11)
    x = rs.rand(500, 9, 7, 2).astype(np.float32)
    w = O.he_init_weights(SMALL, (9, 7, 2), 3, bias_scale=0.1)
    model = _model(nb, SMALL, (9, 7, 2), w, feature_layer=len(SMALL) - 2)
    expr = Expr(k=25, B=100, lambda_=0., batch_size=128)
    expr.pool_images = x
    q = nb.NNAL.CNN_query(model, expr, np.arange(500), 'entropy', None)
    qo, H, post = O.query_entropy_whole(SMALL, w, x, 25)
    eng = nb.get_engine()
    assert np.abs(eng.pool_posteriors() - post).max() < POST_TOL
    assert_topk_equivalent(q, -H, 25, 1e-3 * np.abs(H).max())


def _pw_setup(n_pool, seed, shape=(40, 36, 6), nclass=2):
    ps = (25, 25, 1)
    imgs = synth_volume(shape, 3, seed)
    padded = pad_imgs(imgs, ps)
    stats = vol_stats(imgs)
    rs = np.random.RandomState(seed + 1)
    pool = rs.choice(int(np.prod(shape)), n_pool, replace=False).astype(np.int64)
    layers = O.pw1_layers(nclass)
    probe = O.normalize_batch_eval(O.get_patches(padded, pool[:64], ps), stats).astype(np.float32)
    w = centered_weights(layers, (25, 25, 3), seed + 2, probe)
    return ps, imgs, padded, stats, pool, layers, w


@pytest.mark.parametrize('tc', [0, 1])
def test_pw1_batch_eval_parity(nb, tc):
    """PW_NN.batch_eval posteriors + feature_layer for the PW1 patch CNN vs the float64 oracle
    (SIMT fp32 kernels and tensor-core kernels)."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(384, 20)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.set_tensor_cores(tc)
    try:
        posts, feats = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, ['posteriors', 'feature_layer'])
    finally:
        eng.set_tensor_cores(1)
    op, of = O.batch_eval(layers, w, padded, pool, ps, 128, stats, ['posteriors', 'feature_layer'])
    assert posts.dtype == np.float64 and posts.shape == (384,) and feats.shape == (4096, 384)
    assert (op > .5).any() and (op < .5).any()
    assert np.abs(posts - op).max() < POST_TOL
    assert np.abs(feats - of).max() < 1e-3 * max(1., np.abs(of).max())


def test_pw_entropy_query_single(nb):
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(700, 30)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    expr = Expr(k=50, B=200, lambda_=0., patch_shape=ps, ntb=256, stats=stats)
    q = nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
    qo, posts = O.query_entropy_single(layers, w, padded, pool, ps, 256, stats, 50)
    assert q.shape == (50,)
    assert_topk_equivalent(q, np.abs(posts - .5), 50, POST_TOL)


def test_pw_entropy_query_multimg(nb):
    ps = (25, 25, 1)
    S, m = 3, 3
    shape = (34, 30, 4)
    allp, pools, st = [], [], np.zeros((S, 2 * m))
    rs = np.random.RandomState(41)
    for s in range(S):
        imgs = synth_volume(shape, m, 50 + s)
        allp.append(pad_imgs(imgs, ps) + [(rs.rand(*shape) > .5).astype(np.int8)])
        pools.append(list(rs.choice(int(np.prod(shape)), [150, 0, 230][s], replace=False)))
        for j in range(m):
            st[s, 2 * j], st[s, 2 * j + 1] = imgs[j].mean(), imgs[j].std()
    layers = O.pw1_layers(2)
    probe = O.normalize_batch_eval(O.get_patches(allp[0][:m], pools[0][:64], ps),
                                   [[st[0, 2 * j], st[0, 2 * j + 1]] for j in range(m)]).astype(np.float32)
    w = centered_weights(layers, (25, 25, 3), 60, probe)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    expr = Expr(k=40, B=100, lambda_=0., patch_shape=ps, ntb=128)
    expr.train_stats = st
    Q = nb.PW_NNAL.query_multimg(expr, model, None, allp, pools, None, 'entropy')
    sel_o, posts_o = O.bin_uncertainty_filter_multimg(layers, w, allp, pools, ps, 128, st, 40)
    assert len(Q) == S and len(Q[1]) == 0 and sum(len(a) for a in Q) == 40
    # compare as global positions with tie tolerance
    sizes = [len(p) for p in pools]
    offs = np.concatenate([[0], np.cumsum(sizes)])
    got_global = np.concatenate([np.asarray(Q[s]) + offs[s] for s in range(S)])
    all_posts = np.concatenate([O.batch_eval(layers, w, allp[s][:m], pools[s], ps, 128,
                                             [[st[s, 2 * j], st[s, 2 * j + 1]] for j in range(m)], 'posteriors')[0]
                                if sizes[s] else np.zeros(0) for s in range(S)])
    score = np.abs(all_posts - .5)
    kth = np.sort(score)[39]
    assert np.all(score[got_global] <= kth + POST_TOL)
    assert np.all(np.isin(np.where(score < kth - POST_TOL)[0], got_global))


# ------------------------------------------------------------------ tensor-core FC GEMM in isolation
@pytest.mark.parametrize('M,N,K', [(128, 256, 64), (300, 256, 128), (1000, 320, 200), (777, 4096, 4704),
                                   (8192, 512, 192), (50, 64, 4096)])
def test_fc_tcgen05_vs_f64(nb, M, N, K):
    rs = np.random.RandomState(M + N + K)
    A = np.maximum(rs.randn(M, K), 0).astype(np.float32) * 3           # post-ReLU-like activations
    W = (rs.randn(N, K) * np.sqrt(2. / K)).astype(np.float32)
    b = (rs.randn(N) * .1).astype(np.float32)
    ref = np.maximum(A.astype(np.float64) @ W.astype(np.float64).T + b, 0)
    eng = nb.get_engine()
    got_tc = eng.debug_fc(A, W, b, 1, 1)
    got_simt = eng.debug_fc(A, W, b, 1, 0)
    scale = np.abs(ref).max()
    assert np.abs(got_simt - ref).max() < 2e-5 * scale
    assert np.abs(got_tc - ref).max() < 2e-5 * scale, 'bf16x3 tcgen05 GEMM off by %g' % (np.abs(got_tc - ref).max() / scale)


@pytest.mark.parametrize('H,Cin,Cout,ks', [(25, 3, 24, 5), (25, 24, 32, 5), (13, 32, 48, 3), (13, 48, 96, 3)])
@pytest.mark.parametrize('n', [1, 5, 300])
def test_conv_tcgen05_vs_f64(nb, H, Cin, Cout, ks, n):
    """tcgen05 shift-implicit-GEMM conv (PW1 conv2/conv3/conv4 shapes) vs the float64 oracle."""
    rs = np.random.RandomState(H + Cin + n)
    x = np.maximum(rs.randn(n, H, H, Cin), 0).astype(np.float32)
    W = (rs.randn(ks, ks, Cin, Cout) * np.sqrt(2. / (ks * ks * Cin))).astype(np.float32)
    b = (rs.randn(Cout) * .1).astype(np.float32)
    ref = np.maximum(O.conv2d_same(x.astype(np.float64), W.astype(np.float64), b.astype(np.float64)), 0)
    eng = nb.get_engine()
    got_simt = eng.debug_conv(x, W, b, 0)
    scale = np.abs(ref).max()
    assert np.abs(got_simt - ref).max() < 2e-5 * scale
    got_tc = eng.debug_conv(x, W, b, 1)
    err = np.abs(got_tc - ref).max() / scale
    assert err < 3e-5, 'tcgen05 conv off by %g (relative to max)' % err


@pytest.mark.parametrize('H,Cin,Cout,ks', [(25, 24, 32, 5), (25, 3, 24, 5), (13, 32, 48, 3)])
@pytest.mark.parametrize('n', [1, 2, 5, 300])
def test_conv_weight_stationary_vs_f64(nb, H, Cin, Cout, ks, n):
    """Weight-stationary tcgen05 conv (conv_wt.cu: weights on M, 256 raster positions on N, tap pairs stacked
    in the weight rows) vs the float64 oracle, PW1 conv1/conv2/conv3 shapes; odd sample counts exercise the
    partly filled last group and the single-raster band hand-over."""
    rs = np.random.RandomState(7 * H + Cin + n)
    x = np.maximum(rs.randn(n, H, H, Cin), 0).astype(np.float32)
    W = (rs.randn(ks, ks, Cin, Cout) * np.sqrt(2. / (ks * ks * Cin))).astype(np.float32)
    b = (rs.randn(Cout) * .1).astype(np.float32)
    ref = np.maximum(O.conv2d_same(x.astype(np.float64), W.astype(np.float64), b.astype(np.float64)), 0)
    scale = np.abs(ref).max()
    got = nb.get_engine().debug_conv(x, W, b, 2)
    err = np.abs(got - ref).max() / scale
    assert err < 3e-5, 'weight-stationary conv off by %g (relative to max)' % err


@pytest.mark.parametrize('n', [1, 4, 300])
def test_conv1_x_im2col_vs_f64(nb, n):
    """PW1 conv1 with the 5 filter columns folded into the channel axis (input x-im2col'd to 16 elements per position,
    5 x 1 filter over 16 channels: 5 K-steps instead of 13) vs the float64 oracle conv."""
    H, Cin, Cout, ks = 25, 3, 24, 5
    rs = np.random.RandomState(100 + n)
    x = rs.randn(n, H, H, Cin).astype(np.float32)                     # conv1 sees signed (normalised) inputs
    W = (rs.randn(ks, ks, Cin, Cout) * np.sqrt(2. / (ks * ks * Cin))).astype(np.float32)
    b = (rs.randn(Cout) * .1).astype(np.float32)
    ref = np.maximum(O.conv2d_same(x.astype(np.float64), W.astype(np.float64), b.astype(np.float64)), 0)
    got = nb.get_engine().debug_conv(x, W, b, 5)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 3e-5, 'x-im2col conv1 off by %g (relative to max)' % err


@pytest.mark.parametrize('H,Cin,Cout,ks,mode', [(25, 24, 32, 5, 3), (13, 48, 96, 3, 4)])
@pytest.mark.parametrize('n', [1, 3, 300])
def test_conv_fused_pool(nb, H, Cin, Cout, ks, mode, n):
    """conv + the following 2x2/s2 SAME max-pool in one kernel (shared-memory atomicMax raster; conv2 on the
    weight-stationary kernel, conv4 on the positions-on-M kernel) vs oracle conv -> oracle max_pool_same
    (ceil mode: the last pooled row/column covers a single input row/column)."""
    rs = np.random.RandomState(11 + n)
    x = np.maximum(rs.randn(n, H, H, Cin), 0).astype(np.float32)
    W = (rs.randn(ks, ks, Cin, Cout) * np.sqrt(2. / (ks * ks * Cin))).astype(np.float32)
    b = (rs.randn(Cout) * .1).astype(np.float32)
    ref = O.max_pool_same(np.maximum(O.conv2d_same(x.astype(np.float64), W.astype(np.float64), b.astype(np.float64)), 0))
    got = nb.get_engine().debug_conv(x, W, b, mode)
    assert got.shape == ref.shape == (n, (H + 1) // 2, (H + 1) // 2, Cout)
    err = np.abs(got - ref).max() / np.abs(ref).max()
    assert err < 3e-5, 'fused conv+pool off by %g (relative to max)' % err


# ------------------------------------------------------------------ fused gather / device-resident paths
def test_fused_gather_matches_unfused(nb):
    """The gather that writes conv1's fp16 hi/lo planes directly must give the same posteriors as the
    fp32 gather + split pass it replaces (bit-identical: same float64 normalisation, same split); the default pool pass
    -- conv1 gathering its own input, in the x-im2col'd form -- agrees with both within accumulation-order noise."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(300, 44)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    eng = nb.get_engine()
    d = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
    try:
        eng.debug_option('no_fused_conv1', 1)
        a = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
        eng.debug_option('no_fused_gather', 1)
        b = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
    finally:
        eng.debug_option('no_fused_gather', 0)
        eng.debug_option('no_fused_conv1', 0)
    assert np.array_equal(a, b)
    assert np.abs(d - a).max() < 2e-5
    want = O.batch_eval(layers, w, padded, pool, ps, 100, stats, 'posteriors')[0]
    assert np.abs(d - want).max() < POST_TOL


def test_conv1_gathering_its_input_equals_x_im2col_path(nb):
    """conv1 with the gather fused in runs the same MMAs on the same operand planes as the stand-alone x-im2col gather + conv1
    (debug options conv_x16 = 1, no_fused_conv1 = 1): with the float64 normalisation of batch_eval (wt_flags = 8) the
    posteriors are bit-identical, ragged last chunk included; the default float32 two-term normalisation (error below the
    2^-22 the fp16 hi/lo operands keep) moves them by less than 2e-6."""
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(777, 46)
    try:
        nb.reset_engine()
        eng = nb.get_engine()
        eng.debug_option('conv_x16', 1)
        eng.debug_option('chunk', 200)
        model = nb.NN.create_PW1(2)
        model.set_weights(w)
        d = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
        eng.debug_option('wt_flags', 8)
        a = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
        eng.debug_option('wt_flags', 0)
        eng.debug_option('no_fused_conv1', 1)
        b = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
    finally:
        nb.reset_engine()
    assert np.array_equal(a, b)
    assert np.abs(d - a).max() < 2e-6
    want = O.batch_eval(layers, w, padded, pool, ps, 100, stats, 'posteriors')[0]
    assert np.abs(d - want).max() < POST_TOL


def test_x_im2col_gather_path_matches(nb):
    """debug option conv_x16 = 1: the gather writes conv1's x-im2col'd planes (16 elements per position) and conv1 runs as a 5 x 1
    filter over 16 channels.  Same posteriors as the default path within accumulation-order noise, and the fused and
    unfused (fp32 gather + x-im2col split) variants of it agree bit for bit."""
    import os
    ps, imgs, padded, stats, pool, layers, w = _pw_setup(300, 45)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    ref = nb.PW_NN.batch_eval(model, None, padded, pool, ps, 100, stats, 'posteriors')[0]
    try:
        nb.reset_engine()
        nb.get_engine().debug_option('conv_x16', 1)          # before the weights are uploaded: conv1's packed form differs
        nb.get_engine().debug_option('no_fused_conv1', 1)    # (the default pool pass lets conv1 gather its own input)
        model2 = nb.NN.create_PW1(2)
        model2.set_weights(w)
        a = nb.PW_NN.batch_eval(model2, None, padded, pool, ps, 100, stats, 'posteriors')[0]
        nb.get_engine().debug_option('no_fused_gather', 1)
        b = nb.PW_NN.batch_eval(model2, None, padded, pool, ps, 100, stats, 'posteriors')[0]
    finally:
        nb.reset_engine()
    assert np.array_equal(a, b)
    assert np.abs(a - ref).max() < 2e-5
    want = O.batch_eval(layers, w, padded, pool, ps, 100, stats, 'posteriors')[0]
    assert np.abs(a - want).max() < POST_TOL


def test_device_gather_and_entropy_map(nb):
    import torch
    ps = (25, 25, 1)
    imgs = synth_volume((30, 28, 5), 3, 3)
    padded = pad_imgs(imgs, ps)
    stats = vol_stats(imgs)
    eng = nb.get_engine()
    eng.upload(0, padded)
    inds = np.random.RandomState(0).choice(30 * 28 * 5, 500, replace=False).astype(np.int64)
    d_inds = torch.from_numpy(inds).cuda()
    out = torch.empty((500, 25, 25, 3), dtype=torch.float32, device='cuda')
    for norm in (0, 1):
        eng.gather_device(0, d_inds.data_ptr(), 500, ps, np.array(stats), norm, out.data_ptr())
        eng.synchronize()
        ref = O.get_patches(padded, inds, ps)
        if norm:
            ref = O.normalize_batch_eval(ref, stats)
        assert np.array_equal(out.cpu().numpy(), ref.astype(np.float32))      # float32(float64 arithmetic): bit-exact
    for n in (1000, 1003):
        rs = np.random.RandomState(n)
        z = rs.randn(3, n)
        p = (np.exp(z) / np.exp(z).sum(0)).astype(np.float32)
        p[0, 5] = 0.
        dp = torch.from_numpy(p).cuda()
        H = torch.empty(n, dtype=torch.float32, device='cuda')
        eng.entropy_device(dp.data_ptr(), 3, n, 10e-8, H.data_ptr())
        eng.synchronize()
        ref = O.compute_entropy(p.astype(np.float64).copy())
        assert np.allclose(H.cpu().numpy(), ref, rtol=1e-3, atol=1e-7)        # north star: entropy within 1e-3 relative


def test_normalisation_division_is_ieee(nb):
    """Markstein's correctly-rounded division in the gather == numpy's float64 division, incl. awkward sigmas."""
    rs = np.random.RandomState(9)
    imgs = [(rs.rand(20, 18, 3) * 1000).astype(np.float32) for _ in range(3)]
    padded = pad_imgs(imgs, (5, 5, 1))
    S = np.zeros((1, 6))
    for j, sg in enumerate([1. / 3., 30.000000000000004, np.nextafter(2.0, 0)]):       # last: all-ones significand
        S[0, 2 * j], S[0, 2 * j + 1] = 100. / 7. * (j + 1), sg
    allp = [padded + [np.zeros((20, 18, 3), np.int8)]]
    inds = [list(rs.choice(20 * 18 * 3, 400, replace=False))]
    got, _ = nb.patch_utils.get_patches_multimg(allp, inds, (5, 5, 1), S)
    ref, _ = O.get_patches_multimg(allp, inds, (5, 5, 1), S)
    assert np.array_equal(got[0], ref[0])
