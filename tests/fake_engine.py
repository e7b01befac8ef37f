"""NumPy stand-in for ``nnal_b200.engine.Engine`` used by the CPU-only multi-rank tests (gloo,
world_size 2): same method surface as the C-ABI wrapper, arithmetic through the float64 oracle.
TEST INFRASTRUCTURE ONLY -- it lets the host-side logic (pool sharding, top-k merge, the FI greedy
step protocol, index bookkeeping) run without a GPU; the product never imports it."""
import numpy as np

import oracle as O


class FakeEngine(object):
    def __init__(self):
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.vols = {}
        self._m = {}

    one_pass = True          # tests flip this to exercise the two-pass candidate evaluation as well

    def factor_budget_bytes(self):
        return 1 << 60 if self.one_pass else 0

    @property
    def prev_dim(self):
        return self._prev_dim()

    # -- model / volumes ---------------------------------------------------
    def set_model(self, model, sess=None):
        self.layers = [(k, v) for k, v in model.layer_dict.items()]
        self.weights = model.get_weights(sess)
        self.n_class = int(self.layers[-1][1][0])
        self.feature_layer = model.feature_layer_index if model.feature_layer_index is not None else len(self.layers) - 2

    def upload(self, subject, imgs, pads=(0, 0, 0), shared=False):
        self.vols[subject] = [np.asarray(a) for a in imgs]
        self._m[subject] = len(imgs)

    # -- pool pass -----------------------------------------------------------
    def pool_begin(self, n, keep=0):
        self._pool_n = int(n)
        self.keep = keep
        self.post = np.zeros((self.n_class, n))
        if self._mc is not None:
            self.av_post, self.av_ent = np.zeros(n), np.zeros(n)
        self.feat = None
        self.prev = None
        self.score = None

    # -- MC-dropout / committee (same surface as Engine) ------------------------
    dropout_seed = 0
    dropout_pass = 0
    _mc = None                # (T, keep, layers, pos0, first_pass) while MC mode is on

    def set_dropout_seed(self, seed, first_pass=0):
        self.dropout_seed = int(seed)
        self.dropout_pass = int(first_pass)

    def pool_mc_config(self, T, keep_prob, layers, pos0=0):
        first = self.dropout_pass
        if T > 0:
            self._mc = (int(T), float(keep_prob), list(layers), int(pos0), first)
            self.dropout_pass += int(T)
        else:
            self._mc = None
        return first

    def pool_mc_means(self):
        return self.av_post.copy(), self.av_ent.copy()

    def pool_ensemble_accumulate(self, t):
        from oracle import mc_oracle as M
        p = self.post[1].astype(np.float32).astype(np.float64)
        if t == 0:
            self.av_post, self.av_ent = np.zeros(self._pool_n), np.zeros(self._pool_n)
        self.av_post = (p + t * self.av_post) / (t + 1)
        self.av_ent = (M.binary_entropy_bumped(p) + t * self.av_ent) / (t + 1)

    def pool_ensemble_end(self):
        pass

    def _forward_mc(self, x, offset):
        from oracle import mc_oracle as M
        T, keep, layers, pos0, first = self._mc
        n = x.shape[0]
        pos = pos0 + offset + np.arange(n)
        passes = [M.forward_dropout(self.layers, self.weights, x, pos, keep, layers, self.dropout_seed, first + t)[0]
                  .astype(np.float32).astype(np.float64) for t in range(T)]
        av_p, av_e = M.mc_running_means(passes)
        self.av_post[offset:offset + n] = av_p
        self.av_ent[offset:offset + n] = av_e
        self.post[1, offset:offset + n] = passes[-1]
        self.post[0, offset:offset + n] = 1 - passes[-1]

    def _forward(self, x, offset):
        if self._mc is not None:
            return self._forward_mc(x, offset)
        r = O.forward(self.layers, self.weights, x, self.feature_layer, keep_acts=True)
        n = x.shape[0]
        self.post[:, offset:offset + n] = r['posteriors']
        if self.keep >= 1:
            if self.feat is None:
                self.feat = np.zeros((self._pool_n, r['feature_layer'].shape[0]))
            self.feat[offset:offset + n] = r['feature_layer'].T
        if self.keep >= 2:
            a = r['acts'][self.feature_layer]['in']
            if self.prev is None:
                self.prev = np.zeros((self._pool_n, a.shape[0]))
            self.prev[offset:offset + n] = a.T

    def pool_eval(self, subject, inds, offset, patch_shape, stats, norm_mode=1, shape=None):
        inds = np.asarray(inds)
        if len(inds) == 0:
            return
        x = O.get_patches(self.vols[subject], inds, patch_shape)
        x = O.normalize_batch_eval(x, [list(s) for s in np.asarray(stats)]).astype(np.float32)
        self._forward(x, offset)

    def pool_eval_images(self, x, offset):
        if len(x):
            self._forward(np.asarray(x, dtype=np.float32), offset)

    def pool_posteriors(self):
        return self.post.astype(np.float32)

    def pool_features(self, start=0, n=None):
        n = self._pool_n - start if n is None else n
        return self.feat[start:start + n].T.astype(np.float32)

    def pool_score(self, kind, eps=0.0):
        p = self.post
        if kind in (10, 11):
            from oracle import mc_oracle as M
            self.score = M.mc_entropy_scores(self.av_post) if kind == 10 else -M.bald_scores(self.av_post, self.av_ent)
        elif kind == 0:
            self.score = np.abs(p[1] - .5)
        elif kind in (1, 2):
            q = np.where(p == 0, eps, p)
            h = np.sum(q * np.log(q), axis=0)
            self.score = h if kind == 1 else -h
        else:
            self.score = -O.fi_trace_score(p, self.feat.T)

    def pool_topk(self, k, with_scores=False):
        k = int(min(max(k, 0), self._pool_n))
        idx = O.stable_topk(self.score, k).astype(np.int64)
        return (idx, self.score[idx]) if with_scores else idx

    # -- representativeness queries ------------------------------------------------
    @property
    def feat_dim(self):
        return self._feat_dim()

    def pool_feature_rows(self, pos):
        pos = np.asarray(pos, dtype=np.int64)
        if self.feat is None:
            return np.zeros((0, self._feat_dim()), dtype=np.float32)
        return self.feat[pos].astype(np.float32)

    def _unit_rows(self):
        F = (self.feat if self.feat is not None else np.zeros((0, self._feat_dim()))).astype(np.float32).astype(np.float64)
        nr = np.sqrt(np.sum(F ** 2, axis=1))
        self._inorm = np.where(nr > 0, 1. / np.where(nr > 0, nr, 1.), 0.)
        return F * self._inorm[:, None]

    def rep_set(self, cols, excl_pos, k):
        U = self._unit_rows()
        Cn = np.asarray(cols, dtype=np.float64)
        cn = np.sqrt(np.sum(Cn ** 2, axis=1))
        Cn = Cn * np.where(cn > 0, 1. / np.where(cn > 0, cn, 1.), 0.)[:, None]
        self.rep_S = U @ Cn.T
        self.rep_mask = self._inorm > 0
        self.rep_mask[np.asarray(excl_pos, dtype=np.int64)] = False
        self.rep_cur = np.full(len(U), -np.inf)
        self.rep_taken = np.zeros(len(Cn), dtype=bool)
        self._sel, self._val = [], []
        self._rep_B = len(Cn)

    def rep_step_scores(self, ptr):
        B = self._rep_B
        out = self._mem(ptr, 8 * max(B, 1)).view(np.float64)
        sc = np.sum(np.maximum(self.rep_cur[self.rep_mask, None], self.rep_S[self.rep_mask]), axis=0) if B else np.zeros(0)
        sc = np.where(self.rep_taken, -np.inf, sc)
        out[:B] = sc

    def rep_step_pick(self, t, ptr):
        B = self._rep_B
        sc = self._mem(ptr, 8 * max(B, 1)).view(np.float64)[:B].copy()
        sc[self.rep_taken] = -np.inf
        j = int(np.argmax(sc))
        self._sel.append(j)
        self._val.append(sc[j])
        self.rep_taken[j] = True
        self.rep_cur = np.maximum(self.rep_cur, self.rep_S[:, j])

    def rep_greedy(self, k):
        import ctypes
        buf = np.zeros(max(self._rep_B, 1), dtype=np.float64)
        for t in range(int(min(k, self._rep_B))):
            self.rep_step_scores(buf.ctypes.data)
            self.rep_step_pick(t, buf.ctypes.data)
        return self.sel_result(len(self._sel))

    def sel_result(self, k):
        return np.array(self._sel[:k], dtype=np.int64), np.array(self._val[:k], dtype=np.float64)

    def cross_sims(self, F2):
        U = self._unit_rows()
        T = np.asarray(F2, dtype=np.float64)
        tn = np.sqrt(np.sum(T ** 2, axis=1))
        T = T * np.where(tn > 0, 1. / np.where(tn > 0, tn, 1.), 0.)[:, None]
        self.cs = np.where(self._inorm > 0, np.max(U @ T.T, axis=1), np.inf) if len(U) else np.zeros(0)
        return self.cs.copy()

    def cs_begin(self, init, sims0, gids, k):
        U = self._unit_rows()
        self.cs_U = U
        if init == 0:
            self.cs = np.where(self._inorm > 0, -np.inf, np.inf)
        elif init == 1:
            self.cs = np.where(self._inorm > 0, np.asarray(sims0, dtype=np.float64), np.inf)
        self.cs_gids = np.arange(len(U)) if gids is None else np.asarray(gids, dtype=np.int64)
        self._sel, self._val = [], []

    def cs_msg_bytes(self):
        return (32 + 4 * self._feat_dim() + 15) // 16 * 16

    def cs_step_pack(self, t, ptr):
        buf = self._mem(ptr, self.cs_msg_bytes())
        hdr = buf[:32].view(np.float64)
        if len(self.cs) == 0 or np.min(self.cs) == np.inf:
            hdr[0] = np.inf
            buf[8:16].view(np.int64)[0] = np.iinfo(np.int64).max
            self._cand = -1
            return
        q = int(np.argmin(self.cs))
        self._cand = q
        hdr[0] = self.cs[q]
        buf[8:16].view(np.int64)[0] = self.cs_gids[q]
        hdr[2] = self._inorm[q]
        buf[32:32 + 4 * self._feat_dim()].view(np.float32)[:] = self.feat[q].astype(np.float32)

    def cs_step_apply_gathered(self, t, ptr, world, rank):
        nb = self.cs_msg_bytes()
        buf = self._mem(ptr, nb * world)
        best, bv, bg = 0, np.inf, np.iinfo(np.int64).max
        for r in range(world):
            m = buf[r * nb:(r + 1) * nb]
            v, g = m[:8].view(np.float64)[0], m[8:16].view(np.int64)[0]
            if v < bv or (v == bv and g < bg):
                best, bv, bg = r, v, g
        m = buf[best * nb:(best + 1) * nb]
        inorm = m[16:24].view(np.float64)[0]
        f = m[32:32 + 4 * self._feat_dim()].view(np.float32).astype(np.float64)
        if best == rank and np.isfinite(bv):
            self.cs[self._cand] = np.inf
        if len(self.cs):
            raw = self.feat.astype(np.float32).astype(np.float64)
            s_ind = (raw @ f) * self._inorm * inorm
            live = self.cs < np.inf
            self.cs[live] = np.maximum(self.cs[live], s_ind[live])
        self._sel.append(int(bg) if np.isfinite(bv) else -1)
        self._val.append(bv)

    def cs_greedy(self, k):
        buf = np.zeros(self.cs_msg_bytes(), dtype=np.uint8)
        for t in range(int(min(k, len(self.cs)))):
            self.cs_step_pack(t, buf.ctypes.data)
            self.cs_step_apply_gathered(t, buf.ctypes.data, 1, 0)
        return self.sel_result(len(self._sel))

    # -- Fisher information --------------------------------------------------
    def fi_set_candidates(self, cand=None, n_layers=2):
        rows = np.arange(self._pool_n) if cand is None else np.asarray(cand, dtype=np.int64)
        self.fi_nl = n_layers
        self.fi_p1 = self.post[1, rows]
        f32 = lambda a: a.astype(np.float32).astype(np.float64)       # factors are float32 on the device
        self.fi_U = f32(self.feat[rows]) if len(rows) else np.zeros((0, self._feat_dim()))
        self.fi_A = (f32(self.prev[rows]) if len(rows) else np.zeros((0, self._prev_dim()))) if n_layers == 2 else None
        self.fi_Wl = self.weights[self.layers[-1][0]][0].astype(np.float64)
        self._fi_setup()

    def _feat_dim(self):
        return int(self.layers[self.feature_layer][1][0])

    def _prev_dim(self):
        return int(self.weights[self.layers[self.feature_layer][0]][0].shape[1])

    def _fi_setup(self):
        self.fi_w = self.fi_p1 * (1 - self.fi_p1)
        self.fi_beta = self.fi_Wl[0] - self.fi_Wl[1]
        n = len(self.fi_p1)
        self.fi_diag = np.array([self._kern(i, self.fi_U[i], None if self.fi_A is None else self.fi_A[i],
                                            np.sqrt(self.fi_w[i])) for i in range(n)]) if n else np.zeros(0)

    def _kern(self, i, u, a, sw):
        k = 2. * (self.fi_U[i] @ u + 1.)
        if self.fi_nl == 2:
            mk = ((self.fi_U[i] > 0) & (u > 0)) @ (self.fi_beta ** 2)
            k += mk * (self.fi_A[i] @ a + 1.)
        return np.sqrt(self.fi_w[i]) * sw * k

    def fi_info(self):
        d = self.fi_U.shape[1] if self.fi_U.ndim == 2 and self.fi_U.shape[1] else self._feat_dim()
        dp = self._prev_dim() if self.fi_nl == 2 else 0
        return {'n': len(self.fi_p1), 'n_layers': self.fi_nl, 'd': d, 'd_prev': dp,
                'D': O.last_layers_dim(2, d, dp if self.fi_nl == 2 else None)}

    def fi_begin(self, k, delta):
        n = len(self.fi_p1)
        self.fi_delta = delta
        self.fi_avail = np.ones(n, dtype=bool)
        self.fi_kcols = np.zeros((k, n))
        self.fi_kss = np.zeros((k, k))

    def fi_step_local_best(self, t):
        n = len(self.fi_p1)
        alpha = (t + 1) * self.fi_delta
        trc = 0.
        if t > 0:
            C = np.linalg.inv(alpha * np.eye(t) + self.fi_kss[:t, :t])
            trc = float(np.trace(C))
        if n == 0:
            return float('inf'), -1, trc
        if t == 0:
            r, e = self.fi_diag.copy(), np.zeros(n)
        else:
            kj = self.fi_kcols[:t].T
            Y = kj @ C
            r = self.fi_diag - np.sum(Y * kj, axis=1)
            e = np.sum(Y * Y, axis=1)
        loss = (1. + e) / (alpha + r)
        loss[~self.fi_avail] = np.inf
        j = int(np.argmin(loss))
        if not np.isfinite(loss[j]):
            return float('inf'), -1, trc
        return float(loss[j]), j, trc

    def fi_factor_len(self, step):
        i = self.fi_info()
        return i['d'] + i['d_prev'] + 2 + 2 * (step + 1)

    def fi_winner_factors(self, step, cand):
        row = np.append(self.fi_kcols[:step, cand], self.fi_diag[cand])
        parts = [self.fi_U[cand].astype(np.float32)]
        if self.fi_nl == 2:
            parts.append(self.fi_A[cand].astype(np.float32))
        parts.append(np.array([np.sqrt(self.fi_w[cand])], dtype=np.float64).view(np.float32))
        parts.append(row.astype(np.float64).view(np.float32))
        return np.concatenate(parts)

    def fi_step_apply(self, t, f, owner_is_local, cand_local):
        i = self.fi_info()
        d, dp = i['d'], i['d_prev']
        u = f[:d].astype(np.float64)
        a = f[d:d + dp].astype(np.float64) if dp else None
        sw = float(np.ascontiguousarray(f[d + dp:d + dp + 2]).view(np.float64)[0])
        row = np.ascontiguousarray(f[d + dp + 2:]).view(np.float64)
        self.fi_kss[t, :t + 1] = row
        self.fi_kss[:t + 1, t] = row
        if owner_is_local:
            self.fi_avail[cand_local] = False
        for j in range(len(self.fi_p1)):
            self.fi_kcols[t, j] = self._kern(j, u, a, sw)

    # -- device-resident step protocol, emulated on host memory -----------------------------------
    message_device = 'cpu'
    stream = None

    def fi_set_gids(self, gids):
        self.fi_gids = np.asarray(gids, dtype=np.int64)

    def fi_msg_bytes(self):
        i = self.fi_info()
        return (32 + 8 * self.fi_kcols.shape[0] + 4 * (i['d'] + i['d_prev']) + 15) // 16 * 16

    @staticmethod
    def _mem(ptr, nbytes):
        import ctypes
        return np.ctypeslib.as_array((ctypes.c_uint8 * nbytes).from_address(ptr))

    def fi_step_pack(self, t, ptr):
        nb = self.fi_msg_bytes()
        buf = self._mem(ptr, nb)
        i = self.fi_info()
        kcap = self.fi_kcols.shape[0]
        loss, cand, trc = self.fi_step_local_best(t)
        self._trc, self._cand = trc, cand
        hdr = buf[:32].view(np.float64)
        if cand < 0:
            hdr[0] = np.inf
            buf[8:16].view(np.int64)[0] = np.iinfo(np.int64).max
            return
        hdr[0] = loss
        buf[8:16].view(np.int64)[0] = self.fi_gids[cand] if getattr(self, 'fi_gids', None) is not None and len(self.fi_gids) else cand
        hdr[2] = np.sqrt(self.fi_w[cand])
        row = buf[32:32 + 8 * kcap].view(np.float64)
        row[:t] = self.fi_kcols[:t, cand]
        row[t] = self.fi_diag[cand]
        f = buf[32 + 8 * kcap:32 + 8 * kcap + 4 * (i['d'] + i['d_prev'])].view(np.float32)
        f[:i['d']] = self.fi_U[cand]
        if i['d_prev']:
            f[i['d']:] = self.fi_A[cand]

    def fi_step_apply_gathered(self, t, ptr, world, rank):
        nb = self.fi_msg_bytes()
        buf = self._mem(ptr, nb * world)
        i = self.fi_info()
        kcap = self.fi_kcols.shape[0]
        best, bl, bg = 0, np.inf, np.iinfo(np.int64).max
        for r in range(world):
            m = buf[r * nb:(r + 1) * nb]
            l, g = m[:8].view(np.float64)[0], m[8:16].view(np.int64)[0]
            if l < bl or (l == bl and g < bg):
                best, bl, bg = r, l, g
        m = buf[best * nb:(best + 1) * nb]
        sw = m[16:24].view(np.float64)[0]
        row = m[32:32 + 8 * kcap].view(np.float64)[:t + 1].copy()
        f = m[32 + 8 * kcap:32 + 8 * kcap + 4 * (i['d'] + i['d_prev'])].view(np.float32)
        u = f[:i['d']].astype(np.float64)
        a = f[i['d']:].astype(np.float64) if i['d_prev'] else None
        self.fi_kss[t, :t + 1] = row
        self.fi_kss[:t + 1, t] = row
        if best == rank and np.isfinite(bl):
            self.fi_avail[self._cand] = False
        for j in range(len(self.fi_p1)):
            self.fi_kcols[t, j] = self._kern(j, u, a, sw)
        if not hasattr(self, '_sel') or t == 0:
            self._sel, self._red = [], []
        self._sel.append(int(bg) if np.isfinite(bl) else -1)
        self._red.append((t + 1) * (self._trc + bl))

    def fi_result(self, k):
        return np.array(self._sel[:k], dtype=np.int64), np.array(self._red[:k])

    # -- weighted Gram / primal objective (float64 through the oracle) ------------------------------------
    def _gram(self, idx, q):
        d = self.fi_info()['d']
        Ut = np.concatenate([self.fi_U[idx], np.ones((len(idx), 1))], axis=1) if len(idx) else np.zeros((0, d + 1))
        wq = np.asarray(q, dtype=np.float64) * self.fi_w[idx]
        self.fi_H = (Ut * wq[:, None]).T @ Ut
        return self.fi_H

    def fi_gram(self, q=None, read=True):
        n = len(self.fi_p1)
        H = self._gram(np.arange(n), np.full(n, 1. / max(n, 1)) if q is None else q)
        return H.astype(np.float32) if read else None

    def fi_gram_subset(self, cand, q_sub, read=False):
        H = self._gram(np.asarray(cand, dtype=np.int64), q_sub)
        return H.astype(np.float32) if read else None

    def fi_gram_read(self):
        return self.fi_H.copy()

    def fi_gram_allreduce_host(self):
        import torch
        import torch.distributed as td
        t = torch.from_numpy(np.ascontiguousarray(self.fi_H))
        td.all_reduce(t)
        self.fi_H = t.numpy()
        return self.fi_H.size * 4

    def fi_gram_solve(self, delta, scale=2.0, d_G2_ptr=None):
        n = self.fi_H.shape[0]
        Minv = np.linalg.inv(delta * np.eye(n) + scale * self.fi_H)
        ratio = None if d_G2_ptr is None else float(np.sum(Minv * (delta * np.eye(n) + scale * np.asarray(d_G2_ptr))))
        return float(np.trace(Minv)), ratio

    # -- the reference's literal FI pipeline (shrunk gradients + SDP), through the float64 oracle ------------
    def fi_shrunk_tau(self):
        return sum(1 for _, sp in self.layers if sp[1] != 'pool')

    def fi_shrunk_images(self, x):
        post, g = O.shrunk_class_gradients(self.layers, self.weights, np.asarray(x, dtype=np.float32))
        return post.astype(np.float32), g

    def fi_shrunk_voxels(self, subject, inds, patch_shape, stats, norm_mode=1, shape=None):
        x = O.normalize_batch_eval(O.get_patches(self.vols[subject], np.asarray(inds), patch_shape), stats)
        return self.fi_shrunk_images(x.astype(np.float32))

    def sdp_query_distribution(self, A, tol=1e-4, max_iter=200000, gamma=1.0):
        q, t, phi, gap, it = O.sdp_solve(np.asarray(A), tol, max_iter, 0.5)    # the monotone exponent; the device may start faster
        return {'q': q, 't': t, 'objective': phi, 'gap': gap, 'iterations': it}

    def sdp_query_distribution_reg(self, A, lambda_, X, tol=1e-4, max_iter=20000):
        q, t, Phi, gap, it = O.sdp_solve_reg(np.asarray(A), lambda_, np.asarray(X), tol, max_iter)
        return {'q': q, 't': t, 'objective': Phi, 'gap': gap, 'iterations': it}

    def sdp_from_shrunk(self, g, p1, diag_load, tol=1e-4, max_iter=200000, gamma=1.0):
        return self.sdp_query_distribution(O.gen_A_matrices(g[0], g[1], np.asarray(p1, dtype=np.float64), diag_load), tol, max_iter)

    def fi_greedy(self, k, delta):
        k = int(min(k, len(self.fi_p1)))
        self.fi_begin(max(k, 1), delta)
        D = self.fi_info()['D']
        sel, obj, red = [], [], []
        for t in range(k):
            loss, cand, trc = self.fi_step_local_best(t)
            self.fi_step_apply(t, self.fi_winner_factors(t, cand), True, cand)
            sel.append(cand)
            red.append((t + 1) * (trc + loss))
            obj.append((D - (t + 1)) / delta + red[-1])
        return np.array(sel, dtype=np.int64), np.array(obj), np.array(red)
