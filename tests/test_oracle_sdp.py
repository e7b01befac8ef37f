"""CPU tests of the oracle's restatement of the reference's literal FI pipeline (shrunk coordinates + SDP) and of the
host-side A-matrix assembly (no GPU needed)."""
import numpy as np

import oracle as O
from tests.test_gpu_parity import SMALL


def _rand_A(n, tau, delta, seed, scale=0.05):
    rs = np.random.RandomState(seed)
    s = rs.randn(n, tau) * scale
    p = rs.rand(n)
    g = np.stack([p[:, None] * s, -(1 - p)[:, None] * s])
    return O.gen_A_matrices(g[0], g[1], p, delta), g, p


def test_sdp_solve_matches_slsqp_and_certificate():
    A, _, _ = _rand_A(40, 4, 1e-3, 0)
    q, t, phi, gap, it = O.sdp_solve(A, 1e-6)
    q2, phi2 = O.sdp_solve_slsqp(A)
    assert abs(phi / phi2 - 1) < 1e-5
    assert abs(q.sum() - 1) < 1e-12 and np.all(q >= 0)
    assert abs(t.sum() - phi) < 1e-9 * phi and abs(O.sdp_objective(A, q) - phi) < 1e-9 * phi
    phi_c, gap_c = O.sdp_certificate(A, q)
    assert abs(phi_c - phi) < 1e-9 * phi and gap_c <= 1e-6
    # the certificate is a valid bound: no feasible q does better than phi (1 - gap)
    rs = np.random.RandomState(1)
    for _ in range(50):
        qq = rs.dirichlet(np.ones(40) * .2)
        assert O.sdp_objective(A, qq) >= phi * (1 - gap_c) - 1e-9 * phi
    # the SDP's LMIs hold at (q, t): [[M, e_j], [e_j^T, t_j]] >= 0 (NNAL_tools.py:589-602)
    M = np.tensordot(q, np.array(A), axes=(0, 0))
    for j in range(4):
        e = np.zeros((4, 1)); e[j] = 1
        blk = np.block([[M, e], [e.T, np.array([[t[j]]])]])
        assert np.linalg.eigvalsh(blk).min() > -1e-9 * np.abs(blk).max()


def test_binary_shrunk_gradients_are_multiples_of_one_pass():
    """d log p_0 = p_1 h, d log p_1 = -p_0 h with h = d(z_0 - z_1): what csrc/shrunk.cu's single backward pass for
    binary models relies on (shrink_gradient is linear)."""
    layers = [(n, s) for n, s in SMALL[:-1]] + [('fc3', [2, 'fc'])]
    rs = np.random.RandomState(2)
    x = rs.randn(20, 9, 7, 2).astype(np.float32)
    w = O.he_init_weights(layers, (9, 7, 2), 3, bias_scale=0.1)
    post, g = O.shrunk_class_gradients(layers, w, x)
    s0 = g[0] / post[1][:, None]
    s1 = -g[1] / post[0][:, None]
    assert np.allclose(s0, s1, rtol=1e-9, atol=1e-14)
    assert np.abs(g[:, :, -1]).max() < 1e-14          # last layer: sum_y dz_y = 0


def test_A_assembly_matches_oracle():
    from nnal_b200.PW_NNAL import _A_from_shrunk
    A, g, p = _rand_A(30, 7, 1e-5, 4)
    p[0], p[1] = 1e-9, 1 - 1e-9
    Ao = O.gen_A_matrices(g[0], g[1], p, 1e-5)
    Ah = _A_from_shrunk(g, p, 1e-5)
    assert all(np.array_equal(a, b) for a, b in zip(Ah, Ao))


def test_query_fi_sdp_single_small_pool():
    """The literal pipeline restated end to end on a tiny PW1 pool: <= k unique positions inside the pre-filtered set."""
    from tests.util import pad_imgs, synth_volume, vol_stats
    ps = (25, 25, 1)
    imgs = synth_volume((12, 10, 2), 3, 1)
    padded = pad_imgs(imgs, ps)
    stats = vol_stats(imgs)
    pool = np.arange(0, 240, 12).astype(np.int64)
    layers = O.pw1_layers(2)
    w = O.he_init_weights(layers, (25, 25, 3), 5, bias_scale=0.05)
    u = np.random.RandomState(0).rand(4)
    q, det = O.query_fi_sdp_single(layers, w, padded, pool, ps, 16, stats, 4, 10, u)
    assert len(q) <= 4 and len(np.unique(q)) == len(q) and np.all(np.isin(q, det['sel']))
    assert det['gap'] <= 1e-4 and len(det['A']) == 10 and det['A'][0].shape == (7, 7)


def test_multiclass_A_assembly_matches_oracle():
    from nnal_b200.fi import _A_multiclass_from_shrunk
    rs = np.random.RandomState(6)
    c, B, tau = 12, 25, 5
    g = rs.randn(c, B, tau) * .01
    post = rs.dirichlet(np.ones(c) * .3, B).T.copy()
    post[:, 0] = 0.; post[3, 0] = 1.                     # one-hot posterior: a single class survives
    Ao = O.gen_A_matrices_multiclass(post.copy(), g)
    Ah = _A_multiclass_from_shrunk(post.copy(), g)
    assert all(np.array_equal(a, b) for a, b in zip(Ah, Ao))


def test_sdp_programme_golden(golden):
    """The oracle's solution is feasible for the constraint matrices built by the reference's own
    SDP_query_distribution / inequality_cvx_matrix (golden, captured from the unmodified functions), the reference
    objective c^T x equals tr((sum q_i A_i)^-1), and no feasible point of that programme built from another q does better."""
    from tests.util import assert_feasible_for_reference_sdp
    A = golden['sdp_A']
    q, t, phi, gap, it = O.sdp_solve(A, 1e-8)
    assert np.allclose(q, golden['sdp_q'], rtol=1e-9, atol=1e-14) and np.allclose(t, golden['sdp_t'], rtol=1e-9)
    obj = assert_feasible_for_reference_sdp(golden, q, t)
    assert abs(obj / phi - 1) < 1e-9 and abs(obj / float(golden['sdp_phi']) - 1) < 1e-9
    rs = np.random.RandomState(3)
    for _ in range(30):
        qq = rs.dirichlet(np.ones(len(q)))
        tt = np.diag(np.linalg.inv(np.tensordot(qq, A, axes=(0, 0))))      # the smallest feasible t for this q
        assert assert_feasible_for_reference_sdp(golden, qq, tt * (1 + 1e-9), tol=1e-7) >= obj * (1 - 1e-7)


def test_gen_A_matrices_golden(golden):
    """Golden output of the UNMODIFIED PW_NNAL.gen_A_matrices (driven by a fake session, see
    oracle/check_against_reference.py) == oracle closed form == the package's host assembly."""
    from nnal_b200.PW_NNAL import _A_from_shrunk
    layers = [('conv1', [4, 'conv', [3, 3]]), ('max1', [[2, 2], 'pool']), ('conv2', [6, 'conv', [3, 3]]),
              ('fc1', [10, 'fc']), ('fc2', [2, 'fc'])]
    w = O.he_init_weights(layers, (7, 7, 2), 13, bias_scale=0.1)
    post, g = O.shrunk_class_gradients(layers, w, golden['genA_x'])
    A = O.gen_A_matrices(g[0], g[1], golden['genA_posts'], 1e-5)
    assert np.allclose(np.array(A), golden['genA_out'], rtol=1e-9, atol=1e-18)
    assert np.array_equal(np.array(_A_from_shrunk(g, golden['genA_posts'], 1e-5)), np.array(A))


def test_sdp_solve_reg_matches_slsqp_and_reference_programme(golden):
    """lambda_ > 0: the oracle's feasible multiplicative solver against SLSQP, and against the equalities / objective vector
    the unmodified reference hands to cvxopt (captured: sdpr_c, sdpr_Aeq, sdpr_beq)."""
    rs = np.random.RandomState(3)
    for n, tau, d, lam in [(20, 3, 6, 0.05), (40, 4, 15, 0.5)]:
        A, _, _ = _rand_A(n, tau, 1e-3, n)
        X = np.maximum(rs.randn(d, n), 0)
        X -= X.mean(axis=1, keepdims=True)
        q, t, Phi, gap, it = O.sdp_solve_reg(A, lam, X, 1e-7)
        q2, Phi2 = O.sdp_solve_reg_slsqp(A, lam, X)
        assert abs(Phi / Phi2 - 1) < 1e-6 and gap <= 1e-7
        assert q.min() >= 0 and abs(q.sum() - 1) < 1e-12 and np.abs(X @ q).max() < 1e-12
    A, X, lam = golden['sdp_A'], golden['sdpr_X'], float(golden['sdpr_lambda'])
    q, t, Phi, gap, it = O.sdp_solve_reg(A, lam, X, 1e-9)
    x = np.concatenate([q, t])
    assert np.abs(golden['sdpr_Aeq'] @ x - golden['sdpr_beq'][:, 0]).max() < 1e-12
    assert abs((golden['sdpr_c'][:, 0] @ x).item() - Phi) < 1e-9 * abs(Phi)
    assert np.allclose(q, golden['sdpr_q'], atol=1e-9) and abs(Phi - float(golden['sdpr_Phi'])) < 1e-9 * abs(Phi)
    # the regulariser and the equalities only ever cost A-optimality: the plain optimum is a lower bound of tr(M^-1)
    q0, t0, phi0, gap0, it0 = O.sdp_solve(A, 1e-8)
    assert t.sum() >= phi0 * (1 - 1e-7)
