"""Drop-in for the hot-path functions of the reference's ``patch_utils`` (same names,
argument order and return conventions), running on the GPU through libnnal_b200."""
import numpy as np

from . import _lib as L
from .engine import get_engine


def _rads(patch_shape):
    return [int((patch_shape[i] - 1) / 2.) for i in range(3)]


def get_patches(imgs, inds, patch_shape, padded=True, mask=None):
    """patch_utils.get_patches (patch_utils.py:1087-1173): float64 ``(b, d1, d2, m*d3)``
    patches around raveled voxel ids of the UNPADDED volume; bit-exact.  With ``mask`` also
    returns ``mask[multinds]`` (:1169-1171)."""
    d1, d2, d3 = patch_shape
    if not (d1 % 2 and d2 % 2 and d3 % 2):
        raise ValueError('could not broadcast input array: patch_shape must be odd')
    eng = get_engine()
    rads = _rads(patch_shape)
    pads = (0, 0, 0) if padded else tuple(rads)
    imgs = list(imgs)
    eng.upload(0, imgs, pads)
    inds = np.asarray(inds)
    out = eng.gather(0, inds, patch_shape, shape=np.asarray(imgs[0]).shape, pads=pads)
    if mask is not None:
        pshape = np.asarray(imgs[0]).shape
        orig = tuple(pshape[i] + 2 * pads[i] - 2 * rads[i] for i in range(3))
        labels = mask[np.unravel_index(inds, orig)]
        return out, labels
    return out


def get_patches_multimg(all_padded_imgs, img_inds, patch_shape, stats):
    """patch_utils.get_patches_multimg (patch_utils.py:1175-1212): per-subject gather with the
    subject's mask (last list element) and float64 normalisation of modality block k with
    ``stats[j,2k], stats[j,2k+1]`` (:1203-1207)."""
    eng = get_engine()
    m = len(all_padded_imgs[0]) - 1
    s = len(img_inds)
    rads = _rads(patch_shape)
    b_patches = [[] for _ in range(s)]
    b_labels = [[] for _ in range(s)]
    for j in range(s):
        if len(img_inds[j]) > 0:
            imgs = list(all_padded_imgs[j][:m])
            eng.upload(j, imgs)
            st = np.array([[stats[j, 2 * k], stats[j, 2 * k + 1]] for k in range(m)], dtype=np.float64)
            inds = np.asarray(img_inds[j])
            b_patches[j] = eng.gather(j, inds, patch_shape, st, L.NORM_MULTIMG, shape=np.asarray(imgs[0]).shape)
            pshape = np.asarray(imgs[0]).shape
            orig = tuple(pshape[i] - 2 * rads[i] for i in range(3))
            b_labels[j] = all_padded_imgs[j][m][np.unravel_index(inds, orig)]
    return b_patches, b_labels


def global2local_inds(batch_inds, set_sizes):
    """patch_utils.global2local_inds (patch_utils.py:829-866): host-side index bookkeeping
    (order-preserving split of global positions into per-set local positions)."""
    cumvols = np.append(-1, np.cumsum(set_sizes) - 1)
    set_inds = cumvols.searchsorted(batch_inds) - 1
    return [np.array(batch_inds)[set_inds == i] - cumvols[i] - 1 for i in range(len(set_sizes))]
