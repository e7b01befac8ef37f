"""CPU tests of the oracle itself: against the golden vectors produced by the
reference's own NumPy helpers (oracle/check_against_reference.py) and against
torch-CPU float64 for the TensorFlow-side semantics (forward + autograd)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle as O


def test_gather_golden(golden):
    for ci in range(4):
        imgs = list(golden['gather%d_imgs' % ci])
        out = O.get_patches(imgs, golden['gather%d_inds' % ci], tuple(golden['gather%d_pshape' % ci]))
        assert out.dtype == np.float64
        assert np.array_equal(out, golden['gather%d_out' % ci])


def test_gather_multimg_golden(golden):
    imgs, masks = golden['multi_imgs'], golden['multi_masks']
    allp = [list(imgs[s]) + [masks[s]] for s in range(imgs.shape[0])]
    inds = [list(golden['multi_inds%d' % s]) for s in range(3)]
    p, l = O.get_patches_multimg(allp, inds, (5, 5, 1), golden['multi_stats'])
    for s in range(3):
        if len(inds[s]) == 0:
            assert p[s] == []
            continue
        assert np.array_equal(p[s], golden['multi_out%d' % s])
        assert np.array_equal(l[s], golden['multi_labels%d' % s])


def test_g2l_entropy_golden(golden):
    out = O.global2local_inds(golden['g2l_inds'], golden['g2l_sizes'])
    for i, a in enumerate(out):
        assert np.array_equal(a, golden['g2l_out%d' % i])
    P = golden['entropy_P'].copy()
    assert np.array_equal(O.compute_entropy(P), golden['entropy_H'])
    assert P[1, 3] == 10e-8                      # in-place bump (NNAL_tools.py:80)
    P = golden['entropy_P'].copy()
    assert np.array_equal(O.uncertainty_filtering(P, 9), golden['unc_sel'])
    assert np.array_equal(O.binary_uncertainty_filter(golden['bin_posts'], 20), golden['bin_sel'])
    assert np.array_equal(O.sample_query_dstr(golden['sample_q'].copy(), 8, golden['sample_u']),
                          golden['sample_out'])


def _torch_forward(layers, weights, x, dtype=torch.float64):
    """Independent torch restatement of the TF semantics."""
    h = torch.as_tensor(x, dtype=dtype).permute(0, 3, 1, 2)       # NCHW
    flat = False
    params = {}
    for i, (name, spec) in enumerate(layers):
        last = i == len(layers) - 1
        if spec[1] == 'conv':
            W = torch.tensor(weights[name][0], dtype=dtype, requires_grad=True)
            b = torch.tensor(weights[name][1], dtype=dtype, requires_grad=True)
            params[name] = (W, b)
            h = F.relu(F.conv2d(h, W.permute(3, 2, 0, 1), b, padding=(W.shape[0] // 2, W.shape[1] // 2)))
        elif spec[1] == 'pool':
            h = F.max_pool2d(h, spec[0][0], spec[0][0], ceil_mode=True)
        else:
            if not flat:
                # NCHW -> flat row c*(W*H) + w*H + h
                h = h.permute(1, 3, 2, 0).reshape(-1, h.shape[0])
                flat = True
            W = torch.tensor(weights[name][0], dtype=dtype, requires_grad=True)
            b = torch.tensor(weights[name][1], dtype=dtype, requires_grad=True)
            params[name] = (W, b)
            h = W @ h + b
            if not last:
                h = F.relu(h)
    return h, params


SMALL = [('conv1', [4, 'conv', [3, 3]]), ('conv2', [5, 'conv', [5, 5]]), ('max1', [[2, 2], 'pool']),
         ('conv3', [6, 'conv', [3, 3]]), ('max2', [[2, 2], 'pool']),
         ('fc1', [16, 'fc']), ('fc2', [12, 'fc']), ('fc3', [3, 'fc'])]


def test_forward_vs_torch_f64():
    rs = np.random.RandomState(0)
    w = O.he_init_weights(SMALL, (9, 7, 2), 5, bias_scale=0.1)
    x = rs.randn(6, 9, 7, 2)
    r = O.forward(SMALL, w, x, feature_layer=len(SMALL) - 2)
    logits, _ = _torch_forward(SMALL, w, x)
    assert np.allclose(r['output'], logits.detach().numpy(), rtol=1e-12, atol=1e-12)
    post = torch.softmax(logits, 0).detach().numpy()
    assert np.allclose(r['posteriors'], post, rtol=1e-12, atol=1e-14)
    assert r['feature_layer'].shape == (12, 6)


def test_pw1_shapes_and_flops():
    layers = O.pw1_layers(2)
    shp = O.layer_shapes(layers, (25, 25, 3))
    assert shp[2] == (13, 13, 32) and shp[5] == (7, 7, 96) and shp[-1] == (2,)
    w = O.he_init_weights(layers, (25, 25, 3), 4)
    assert w['fc1'][0].shape == (4096, 4704) and w['conv2'][0].shape == (5, 5, 24, 32)
    nparams = sum(W.size + b.size for W, b in w.values())
    assert abs(nparams - 36.14e6) < 0.01e6


def test_explicit_gradients_vs_autograd():
    rs = np.random.RandomState(1)
    w = O.he_init_weights(SMALL, (9, 7, 2), 6, bias_scale=0.1)
    x = rs.randn(1, 9, 7, 2)
    names = [n for n, s in SMALL if s[1] != 'pool']
    for y in range(3):
        logits, params = _torch_forward(SMALL, w, x)
        lp = torch.log_softmax(logits, 0)[y, 0]
        flat = [t for n in names for t in params[n]]
        tg = torch.autograd.grad(lp, flat)
        og = O.explicit_class_gradients(SMALL, w, x, y)
        assert len(og) == len(tg)
        for a, b in zip(og, tg):
            assert a.shape == tuple(b.shape)
            assert np.allclose(a, b.numpy(), rtol=1e-9, atol=1e-12)
    # closed-form shrink == shrink of explicit gradients; last layer component ~ 0 (H6)
    post, g = O.shrunk_class_gradients(SMALL, w, x)
    for y in range(3):
        sg = O.shrink_gradient(O.explicit_class_gradients(SMALL, w, x, y))
        assert np.allclose(sg, g[y, 0], rtol=1e-9, atol=1e-15)
        assert abs(g[y, 0, -1]) < 1e-15


def test_shrink_golden(golden):
    layers = [('conv1', [4, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']),
              ('conv2', [6, 'conv', [3, 3]]), ('max2', [[2, 2], 'pool']),
              ('fc1', [16, 'fc']), ('fc2', [12, 'fc']), ('fc3', [3, 'fc'])]
    w = O.he_init_weights(layers, (7, 7, 2), 11, bias_scale=0.1)
    post, g = O.shrunk_class_gradients(layers, w, golden['shrink_x'])
    assert np.allclose(g, golden['shrink_g'], rtol=1e-12, atol=1e-18)
    assert np.allclose(post, golden['shrink_post'], rtol=1e-12)


def test_llfc_and_trace():
    rs = np.random.RandomState(2)
    c, d, n = 4, 6, 5
    P = rs.dirichlet(np.ones(c), size=n).T
    U = rs.randn(d, n)
    for i in range(n):
        H = O.LLFC_hess(P[:, i:i + 1], U[:, i:i + 1])
        # Hessian of log-loss = -FI ; trace identity (NNAL.py:124-139)
        assert np.isclose(-np.trace(H), O.fi_trace_score(P[:, i:i + 1], U[:, i:i + 1])[0])
        # FI = sum_y pi_y s_y s_y^T with s_y = LLFC_grads(label=y)
        Fi = np.zeros_like(H)
        for y in range(c):
            s = O.LLFC_grads(P[:, i:i + 1], U[:, i:i + 1], labels=np.array([y]))
            Fi += P[y, i] * (s @ s.T)
        assert np.allclose(Fi, -H, atol=1e-12)


def test_fc_gradnorms_vs_autograd():
    rs = np.random.RandomState(3)
    layers = [('fc1', [7, 'fc']), ('fc2', [5, 'fc']), ('fc3', [2, 'fc'])]
    w = O.he_init_weights(layers, (1, 1, 6), 3, bias_scale=0.2)
    x = rs.randn(4, 1, 1, 6)
    r = O.forward(layers, w, x, keep_acts=True)
    ins = [a['in'] for a in r['acts']]
    norms = O.FC_gradnorms_batch(r['posteriors'], ins, [w[n][0] for n, _ in layers])
    for n in range(4):
        logits, params = _torch_forward(layers, w, x[n:n + 1])
        J0 = torch.softmax(logits, 0)[0, 0]
        for li, (name, _) in enumerate(layers):
            gW, gb = torch.autograd.grad(J0, params[name], retain_graph=True)
            assert np.isclose(norms[li, n], (gW ** 2).sum().item() + (gb ** 2).sum().item(), rtol=1e-9)


def test_greedy_forms_agree():
    """definition (direct, primal) == dual brute force == incremental rank-1 form."""
    rs = np.random.RandomState(4)
    n, D, k, delta = 40, 6, 9, 1e-3
    G = rs.randn(n, D) * rs.rand(n, 1)
    Abar = [np.outer(g, g) for g in G]
    S0, f0 = O.greedy_fi_direct(Abar, delta, k)
    Kt = G @ G.T
    S1, f1 = O.greedy_fi_dual_bruteforce(Kt, 1, D, delta, k)
    S2, f2, red = O.greedy_fi_rank1(Kt, D, delta, k, return_reduced=True)
    assert np.array_equal(S0, S1) and np.array_equal(S0, S2)
    assert np.allclose(f0, f1, rtol=1e-8) and np.allclose(f0, f2, rtol=1e-8)
    # objective equals the reference SDP objective at q = uniform(S)
    A = [a + delta * np.eye(D) for a in Abar]
    q = np.zeros(n)
    q[S0] = 1. / k
    assert np.isclose(O.sdp_objective(A, q), f0[-1], rtol=1e-9)


def test_last_layer_kernel_and_gram():
    rs = np.random.RandomState(5)
    d, n, delta = 5, 7, 1e-2
    U = np.maximum(rs.randn(d, n), 0)
    p1 = rs.rand(n) * .8 + .1
    P = np.stack([1 - p1, p1])
    Kt = O.last_layers_kernel(p1, U)
    D = O.last_layers_dim(2, d)
    # explicit FI matrices via LLFC_hess
    Abar = [-O.LLFC_hess(P[:, i:i + 1], U[:, i:i + 1]) for i in range(n)]
    S = [0, 3, 4]
    f_direct = O.fi_objective_direct(Abar, S, delta)
    f_dual = O.fi_objective_dual(Kt[np.ix_(S, S)], len(S), D, delta)
    assert np.isclose(f_direct, f_dual, rtol=1e-9)
    wq = np.zeros(n)
    wq[S] = (p1 * (1 - p1))[S] / len(S)
    f_gram = O.fi_objective_from_gram(O.weighted_gram(U, wq), 2, delta)
    assert np.isclose(f_direct, f_gram, rtol=1e-9)


def test_two_layer_kernel_vs_explicit():
    rs = np.random.RandomState(6)
    layers = [('fc1', [6, 'fc']), ('fc2', [5, 'fc']), ('fc3', [2, 'fc'])]
    w = O.he_init_weights(layers, (1, 1, 4), 8, bias_scale=0.3)
    x = rs.randn(6, 1, 1, 4)
    r = O.forward(layers, w, x, keep_acts=True)
    p1 = r['posteriors'][1]
    U = r['acts'][2]['in']          # fc2 output = input of fc3
    Aprev = r['acts'][1]['in']      # fc1 output = input of fc2
    Kt = O.last_layers_kernel(p1, U, Aprev, w['fc3'][0].astype(np.float64))
    # explicit: F_i = sum_y pi_y s_y s_y^T over params of fc2,fc3
    n = 6
    G = []
    for i in range(n):
        Fi = 0
        for y in range(2):
            g = O.explicit_class_gradients(layers, w, x[i:i + 1], y, grad_layers=['fc2', 'fc3'])
            s = np.concatenate([a.ravel() for a in g])
            Fi = Fi + r['posteriors'][y, i] * np.outer(s, s)
        G.append(Fi)
    D = G[0].shape[0]
    assert D == O.last_layers_dim(2, 5, 6)
    S = [1, 2, 5]
    f_direct = O.fi_objective_direct(G, S, 1e-2)
    f_dual = O.fi_objective_dual(Kt[np.ix_(S, S)], 3, D, 1e-2)
    assert np.isclose(f_direct, f_dual, rtol=1e-8)


def test_gen_A_matrices_rank_one_identity():
    """For c=2, A_i - delta I = p(1-p) gbar gbar^T with g0 = p1*gbar, g1 = -p0*gbar."""
    rs = np.random.RandomState(7)
    gbar = rs.randn(5, 3)
    p = np.array([.3, 1e-8, 1 - 1e-9, .5, .9])
    g0, g1 = gbar * p[:, None], -gbar * (1 - p)[:, None]
    A = O.gen_A_matrices(g0, g1, p, 1e-3)
    for i in (0, 3, 4):
        assert np.allclose(A[i] - 1e-3 * np.eye(3), p[i] * (1 - p[i]) * np.outer(gbar[i], gbar[i]))
    assert np.allclose(A[1] - 1e-3 * np.eye(3), np.outer(g0[1], g0[1]))     # clamped p=0
    assert np.allclose(A[2] - 1e-3 * np.eye(3), np.outer(g1[2], g1[2]))     # clamped p=1


def test_greedy_replay_matches_rank1():
    """greedy_fi_replay (the tolerance-aware checker used by the GPU tests) follows greedy_fi_rank1."""
    rs = np.random.RandomState(3)
    n, d, dp = 60, 12, 10
    U = np.maximum(rs.randn(d, n), 0)
    A = np.maximum(rs.randn(dp, n), 0)
    Wl = rs.randn(2, d)
    p1 = rs.rand(n)
    for two in (False, True):
        Kt = O.last_layers_kernel(p1, U, A if two else None, Wl if two else None)
        D = O.last_layers_dim(2, d, dp if two else None)
        S, obj = O.greedy_fi_rank1(Kt, D, 1e-3, 15)
        rep = O.greedy_fi_replay(Kt, D, 1e-3, S)
        assert np.allclose(rep[:, 0], rep[:, 1], rtol=1e-12)
        assert np.allclose(rep[:, 2], obj, rtol=1e-9)
        # definition: f(S) through explicit conditional FIs
        G = np.concatenate([np.kron(np.array([1., -1.])[:, None], np.concatenate([U, np.ones((1, n))])),
                            ] + ([np.stack([np.kron((Wl.T @ np.array([1., -1.])) * (U[:, i] > 0),
                                                    np.append(A[:, i], 1.)) for i in range(n)], axis=1)] if two else []),
                           axis=0) * np.sqrt(p1 * (1 - p1))[None, :]
        Abar = [np.outer(G[:, i], G[:, i]) for i in range(n)]
        assert np.isclose(O.fi_objective_direct(Abar, list(S[:6]), 1e-3), obj[5], rtol=1e-8)


def test_rep_oracle_golden(golden):
    """Similarity helpers and greedy loops of the representativeness queries vs the reference's own outputs."""
    F1, F2 = golden['sims_F1'], golden['sims_F2']
    assert np.allclose(O.get_self_sims(F1), golden['sims_self'], rtol=1e-13)
    assert np.allclose(O.get_cross_sims(F1, F2), golden['sims_cross'], rtol=1e-13)
    assert np.array_equal(O.greedy_facility_location(golden['fl_sims'], 6)[0], golden['fl_Q'])
    rep = O.facility_location_replay(golden['fl_sims'], golden['fl_Q'])
    assert np.allclose(rep[:, 0], rep[:, 1])
    assert np.array_equal(O.kcenter_greedy(F1, golden['sims_cross'], 5)[0], golden['kc_Q'])
    rep = O.kcenter_replay(F1, golden['sims_cross'], golden['kc_Q'])
    assert np.allclose(rep[:, 0], rep[:, 1])
