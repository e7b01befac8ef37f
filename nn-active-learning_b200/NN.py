"""Model container mirroring the parts of the reference's ``NN.CNN`` the query path reads
(NN.py:56-345, 1319-1359): the ordered ``layer_dict``, ``var_dict`` {layer: [W, b]} in the TF
variable layouts, ``feature_layer`` index and dropout metadata.  No TensorFlow graph is built --
the forward pass runs in libnnal_b200."""
from collections import OrderedDict

import numpy as np


class _Placeholder(object):
    """Stands in for a TensorFlow placeholder (``model.keep_prob``, NN.py:33-35): only ever used as a key of the
    reference's ``x_feed_dict = {model.keep_prob: model.dropout_rate}`` (PW_NNAL.py:68-69, 234-235)."""

    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return '<placeholder %s>' % self.name


class CNN(object):
    """``layer_dict``: ordered {name: [out,'conv',[kh,kw]] | [[p,p],'pool'] | [out,'fc']}
    (NN.py:98-108).  ``input_shape`` = (H, W, C) of the NHWC placeholder (NN.py:1339-1345)."""

    def __init__(self, input_shape, layer_dict, name='CNN', feature_layer=None, dropout=None, probes=[]):
        self.input_shape = tuple(int(v) for v in input_shape)
        self.layer_dict = OrderedDict(layer_dict)
        self.name = name
        self.feature_layer_index = feature_layer
        self.layer_type = [spec[1] for spec in self.layer_dict.values()]
        if dropout:
            self.dropout_layers, self.dropout_rate = dropout[0], dropout[1]
        else:
            self.dropout_layers, self.dropout_rate = [], 1.
        self.probes = list(probes)
        self.keep_prob = _Placeholder('keep_prob')
        self.x = _Placeholder('x')             # key of the patches in a reference-style feed_dict ({model.x: patches})
        self.var_dict = {}
        self.grad_layers = []
        self._version = 0
        # validate the way NN.CNN.add_layer does (NN.py:246-249)
        for spec in self.layer_dict.values():
            if spec[1] not in ('conv', 'fc', 'pool'):
                raise ValueError("Layer's type should be either 'fc', 'conv' or 'pool'.")
        last = list(self.layer_dict.values())[-1]
        self.nclass = int(last[0])

    # -- shapes -----------------------------------------------------------
    def weight_shapes(self):
        H, W, C = self.input_shape
        flat = None
        shapes = OrderedDict()
        for name, spec in self.layer_dict.items():
            if spec[1] == 'conv':
                shapes[name] = ((spec[2][0], spec[2][1], C, spec[0]), (spec[0],))
                C = spec[0]
            elif spec[1] == 'pool':
                s = spec[0][0]
                H, W = -(-H // s), -(-W // s)
            else:
                if flat is None:
                    flat = H * W * C
                shapes[name] = ((spec[0], flat), (spec[0], 1))
                flat = spec[0]
        return shapes

    # -- weights ----------------------------------------------------------
    def set_weights(self, weights):
        """``weights``: {layer: (W, b)} in TF layouts -- conv ``[kh,kw,cin,cout]``/``[cout]``
        (NN.py:270-283), fc ``[out,in]``/``[out,1]`` (NN.py:311-320)."""
        shapes = self.weight_shapes()
        for name, (ws, bs) in shapes.items():
            W, b = weights[name]
            W = np.array(W, dtype=np.float32)          # private copies: later edits of the caller's arrays do not leak in
            b = np.array(b, dtype=np.float32)
            if W.shape != ws or b.size != bs[0]:
                raise ValueError('layer %s: expected W%s b%s, got W%s b%s' % (name, ws, bs, W.shape, b.shape))
            W.setflags(write=False)
            b = b.reshape(bs)
            b.setflags(write=False)
            self.var_dict[name] = [W, b]
        self._version += 1

    def weights_token(self):
        """Changes whenever the weights do (``set_weights`` is the only way in, and it stores read-only copies): lets the
        engine skip re-hashing 145 MB of weights on every query."""
        return self._version

    def get_weights(self, sess=None):
        if not self.var_dict:
            raise RuntimeError('model has no weights: call set_weights / load_weights first')
        return {k: (v[0], v[1]) for k, v in self.var_dict.items()}

    def initialize(self, seed=None, bias_scale=0.):
        """He-normal weights ``N(0, 2/fan_in)``, zero biases (NN.py:1430-1470); ``bias_scale`` > 0 draws ``N(0, bias_scale^2)``
        biases instead (synthetic workloads that exercise the bias path)."""
        rs = np.random.RandomState(seed)
        w = {}
        for name, (ws, bs) in self.weight_shapes().items():
            fan_in = ws[0] * ws[1] * ws[2] if len(ws) > 2 else ws[1]
            W = (rs.randn(*ws) * np.sqrt(2. / fan_in)).astype(np.float32)
            w[name] = (W, (rs.randn(*bs) * bias_scale).astype(np.float32))
        self.set_weights(w)

    @staticmethod
    def _is_hdf5(file_path):
        return str(file_path).lower().endswith(('.h5', '.hdf5', '.hdf'))

    def save_weights(self, file_path):
        """NN.save_weights (NN.py:379-396): ``<layer>/Weight`` and ``<layer>/Bias`` -- as HDF5 groups/datasets when the
        path ends in .h5/.hdf5 (written by ``nnal_b200.hdf5``: h5py's default on-disk format), else as the keys of an NPZ."""
        if self._is_hdf5(file_path):
            from . import hdf5
            hdf5.write_weights(file_path, {name: (W, b) for name, (W, b) in self.var_dict.items()})
            return
        d = {}
        for name, (W, b) in self.var_dict.items():
            d[name + '/Weight'] = W
            d[name + '/Bias'] = b
        np.savez(file_path, **d)

    def load_weights(self, file_path):
        """Inverse of ``save_weights``; HDF5 files are the reference's own weight files (NN.py:379-396)."""
        if self._is_hdf5(file_path):
            from . import hdf5
            try:
                w = hdf5.read_weights(file_path)
            except NotImplementedError:
                import h5py                              # layouts outside the built-in reader (chunked, filtered, ...)
                with h5py.File(file_path, 'r') as f:
                    w = {name: (np.array(f[name]['Weight']), np.array(f[name]['Bias'])) for name in self.weight_shapes()}
            missing = [name for name in self.weight_shapes() if name not in w]
            if missing:
                raise KeyError('weight file %s has no group(s) %s' % (file_path, ', '.join(missing)))
            self.set_weights({name: w[name] for name in self.weight_shapes()})
            return
        z = np.load(file_path)
        self.set_weights({name: (z[name + '/Weight'], z[name + '/Bias']) for name in self.weight_shapes()})

    def perform_assign_ops(self, file_path, sess=None):
        """NN.CNN.perform_assign_ops (NN.py:397-419): load the weights saved at ``file_path`` into the model (the
        reference's HDF5 files, or NPZ files with the same ``<layer>/Weight``, ``<layer>/Bias`` keys)."""
        self.load_weights(file_path)

    def get_gradients(self, grad_layers=[]):
        """Records which layers the FI score factors cover (NN.py:621-645)."""
        self.grad_layers = list(grad_layers)


class ReferenceModelAdapter(object):
    """Wraps a live reference ``NN.CNN`` (TensorFlow) so the engine can read its weights the
    way NN.save_weights does (``var.eval()``, NN.py:391-394).  ``layer_dict`` and
    ``input_shape`` must be supplied because the reference object does not keep them."""

    def __init__(self, ref_model, layer_dict, input_shape, feature_layer=None):
        self.ref = ref_model
        self.layer_dict = OrderedDict(layer_dict)
        self.input_shape = tuple(input_shape)
        self.feature_layer_index = feature_layer
        self.dropout_rate = getattr(ref_model, 'dropout_rate', 1.)
        self.dropout_layers = list(getattr(ref_model, 'dropout_layers', []))
        self.keep_prob = getattr(ref_model, 'keep_prob', _Placeholder('keep_prob'))
        self._version = 0

    def refresh(self):
        """Kept for callers of the first release; no longer needed: the engine re-reads the variables on every query and
        re-uploads whenever their content changed (``Engine.set_model``)."""
        self._version += 1

    def get_weights(self, sess=None):
        out = {}
        for name, spec in self.layer_dict.items():
            if spec[1] == 'pool':
                continue
            W, b = self.ref.var_dict[name]
            out[name] = (np.asarray(W.eval(session=sess)), np.asarray(b.eval(session=sess)))
        return out


def pw1_layer_dict(nclass):
    """Layer dictionary of create_PW1 (NN.py:1328-1336)."""
    return OrderedDict([('conv1', [24, 'conv', [5, 5]]), ('conv2', [32, 'conv', [5, 5]]),
                        ('max1', [[2, 2], 'pool']), ('conv3', [48, 'conv', [3, 3]]),
                        ('conv4', [96, 'conv', [3, 3]]), ('max2', [[2, 2], 'pool']),
                        ('fc1', [4096, 'fc']), ('fc2', [4096, 'fc']), ('fc3', [nclass, 'fc'])])


def create_PW1(nclass, dropout_rate=1., learning_rate=None, optimizer_name=None, patch_shape=(25, 25, 3)):
    """NN.create_PW1 (NN.py:1319-1359): same signature; optimiser arguments are accepted and
    ignored (training stays in the reference)."""
    d = pw1_layer_dict(nclass)
    model = CNN((patch_shape[0], patch_shape[1], patch_shape[2]), d, 'PatchWise',
                feature_layer=len(d) - 2, dropout=[[6, 7, 8], dropout_rate], probes=[5])
    model.get_gradients()
    return model


def _last_layer_factors(model, sess, feed_dict):
    """posteriors [c,n] and feature_layer [d,n] (float32, the layouts of ``sess.run(model.posteriors / feature_layer)``,
    NN.py:184-188, 173-176) of the samples fed as ``feed_dict[model.x]`` -- one forward pass on the device."""
    from .engine import get_engine
    x = np.asarray(feed_dict[model.x], dtype=np.float32)
    eng = get_engine()
    eng.set_model(model, sess)
    n = x.shape[0]
    eng.pool_begin(n, 1)
    eng.pool_eval_images(x, 0)
    post = eng.pool_posteriors()
    U = eng.pool_feature_rows(np.arange(n, dtype=np.int64)).T
    return post, np.ascontiguousarray(U)


def LLFC_grads(model, sess, feed_dict, labels=None):
    """NN.LLFC_grads (NN.py:905-955): gradients of the log-loss w.r.t. the last FC layer, ``[(e_y - pi) (x) u ; (e_y - pi)]``
    as a ``((d+1)c, n)`` matrix (rows: W class-major, then the biases).  ``labels`` None: the model's own predictions are used
    and returned as well (:928-931, 951-953).  The forward pass runs on the device; the assembly is the reference's NumPy
    (including its float32 ``pi * u`` product)."""
    pies, U = _last_layer_factors(model, sess, feed_dict)
    c, n = pies.shape
    d = U.shape[0]
    rep_U = np.tile(U, (c, 1))
    pies_dot_U = np.repeat(pies, d, axis=0) * rep_U
    flag = labels is None
    if flag:
        labels = np.argmax(pies, axis=0)
    hot_labels = np.zeros((c, n))
    for j in range(c):
        hot_labels[j, np.asarray(labels) == j] = 1
    dJ_dW = np.repeat(hot_labels, d, axis=0) * rep_U - pies_dot_U
    dJ_db = hot_labels - pies
    G = np.concatenate((dJ_dW, dJ_db), axis=0)
    return (G, labels) if flag else G


def LLFC_hess(model, sess, feed_dict):
    """NN.LLFC_hess (NN.py:874-903): explicit ``((d+1)c)^2`` Hessian of the soft-max loss w.r.t. the last FC layer for ONE
    sample, ``[[kron(A,uu^T), kron(A,u)],[kron(A,u^T), A]]`` with ``A = -(diag pi - pi pi^T)``.  Kept for callers that want
    the matrix (537 MB at PW1); the query / influence paths use its factored form and never build it."""
    pi, u = _last_layer_factors(model, sess, feed_dict)
    d = u.shape[0]
    c = pi.shape[0]
    repM = np.repeat(pi, c, axis=1) - np.eye(c)
    A = np.diag(pi[:, 0]) @ repM.T
    H = np.zeros(((d + 1) * c, (d + 1) * c))
    H[:c * d, :c * d] = np.kron(A, np.outer(u, u))
    H[:c * d, c * d:] = np.kron(A, u)
    H[c * d:, :c * d] = np.kron(A, u.T)
    H[c * d:, c * d:] = A
    return H
