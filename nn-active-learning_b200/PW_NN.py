"""Drop-in for ``PW_NN.batch_eval`` (PW_NN.py:357-539): pool-wide gather -> normalise ->
forward, on the GPU."""
import numpy as np

from . import _lib as L
from .engine import get_engine


def keep_prob_from_feed(model, x_feed_dict):
    """``{model.keep_prob: rate}`` -> rate, ``{}`` -> None; any other feed is outside the replaced path."""
    if len(x_feed_dict) == 0:
        return None
    if len(x_feed_dict) != 1 or list(x_feed_dict.keys())[0] is not model.keep_prob:
        raise NotImplementedError('only {model.keep_prob: rate} can be fed to the replaced path')
    return float(list(x_feed_dict.values())[0])


def batch_eval(model, sess, img_dat, inds, patch_shape, batch_size, stats, varnames, mask=None,
               x_feed_dict={}):
    """Same signature and return convention as the reference: a list with one array per
    requested variable -- ``posteriors`` -> float64 ``[n]`` P(class 1) (PW_NN.py:526-529),
    ``feature_layer`` -> float64 ``[d, n]`` (:459-461, 530-531), ``prediction`` -> ``[n]``.
    ``batch_size`` (``ntb``) only bounds host batching in the reference; results do not
    depend on it.  ``sess`` is unused.  ``x_feed_dict = {model.keep_prob: rate}`` (PW_NNAL.py:68-69) runs ONE
    stochastic pass with tf.nn.dropout semantics on ``model.dropout_layers`` (NN.py:167-171); the masks come from the
    engine's counter-based generator (``nnal_b200.get_engine().set_dropout_seed``), a new pass id per call.  Image
    paths and ``loss``/``hess_vecp`` are outside the replaced path."""
    if not isinstance(varnames, list):
        varnames = [varnames]
    keep_prob = keep_prob_from_feed(model, x_feed_dict)
    if keep_prob is not None and varnames != ['posteriors']:
        raise NotImplementedError('stochastic passes return posteriors only')
    if not isinstance(img_dat[0], np.ndarray):
        raise NotImplementedError('img_dat must hold the padded arrays (nrrd paths are read by the reference)')
    for var in varnames:
        if var not in ('posteriors', 'feature_layer', 'prediction'):
            raise NotImplementedError('variable %s stays in the reference (training-side)' % var)
    eng = get_engine()
    eng.set_model(model, sess)
    eng.upload(0, list(img_dat))
    inds = np.asarray(inds)
    n = len(inds)
    keep = 1 if 'feature_layer' in varnames else 0
    st = np.array([[stats[j][0], stats[j][1]] for j in range(len(img_dat))], dtype=np.float64)
    if keep_prob is not None:
        eng.pool_mc_config(1, keep_prob, model.dropout_layers)
    try:
        eng.pool_begin(n, keep)
        eng.pool_eval(0, inds, 0, patch_shape, st, L.NORM_BATCH_EVAL, shape=img_dat[0].shape)
    finally:
        if keep_prob is not None:
            eng.pool_mc_config(0, 1., [])
    vals_list = []
    post = None
    for var in varnames:
        if var in ('posteriors', 'prediction'):
            if post is None:
                post = eng.pool_posteriors()
            if var == 'posteriors':
                vals_list += [post[1, :].astype(np.float64)]
            else:
                vals_list += [np.argmax(post, axis=0).astype(np.float64)]
        else:
            vals_list += [eng.pool_features().astype(np.float64)]
    return vals_list
