"""Boundary robustness on the GPU (``pytest -m gpu``): the fp16 range guard of the tensor-core path, the
content-hash caches of volumes and weights, and the device-resident top-k merge used by the multi-GPU queries."""
import numpy as np
import pytest

import oracle as O
from tests.util import centered_weights, pad_imgs, synth_volume, vol_stats

pytestmark = pytest.mark.gpu


class Expr(object):
    def __init__(self, **pars):
        self.pars = pars


@pytest.fixture(scope='module')
def nb():
    import nnal_b200
    return nnal_b200


def _setup(seed=3, n_pool=150, shape=(30, 28, 4)):
    ps = (25, 25, 1)
    imgs = synth_volume(shape, 3, seed)
    padded = pad_imgs(imgs, ps)
    stats = vol_stats(imgs)
    pool = np.random.RandomState(seed + 1).choice(int(np.prod(shape)), n_pool, replace=False).astype(np.int64)
    layers = O.pw1_layers(2)
    probe = O.normalize_batch_eval(O.get_patches(padded, pool[:32], ps), stats).astype(np.float32)
    w = centered_weights(layers, (25, 25, 3), seed + 2, probe)
    return ps, padded, stats, pool, layers, w


def test_fp16_overflow_is_an_error_not_a_saturation(nb):
    """Activations beyond 65504 cannot be carried by the fp16 hi/lo operands: the query fails with NnalOverflowError
    (NNAL_ERR_OVERFLOW) instead of returning saturated posteriors; the FP32 CUDA-core path still evaluates the same
    input, and the context stays usable."""
    from nnal_b200 import _lib as L
    ps, padded, stats, pool, layers, w = _setup()
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    big = [p * np.float32(40000.) for p in padded]           # inputs ~1e5 after the (unchanged) normalisation: outside fp16
    expr = Expr(k=10, B=50, lambda_=0., patch_shape=ps, ntb=64, stats=stats)
    with pytest.raises(L.NnalOverflowError):
        nb.PW_NNAL.CNN_query(expr, model, None, big, pool, None, 'entropy')
    with pytest.raises(OverflowError):                       # also an OverflowError for callers that do not know the package
        nb.PW_NN.batch_eval(model, None, big, pool, ps, 64, stats, 'posteriors')
    # the flag is cleared by the failing call: the same context scores a regular input right after
    q = nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
    qo, posts = O.query_entropy_single(layers, w, padded, pool, ps, 64, stats, 10)
    assert np.abs(nb.get_engine().pool_posteriors()[1] - posts).max() < 1e-4
    # FP32 kernels: no fp16 operands, no limit
    eng = nb.get_engine()
    eng.set_tensor_cores(False)
    try:
        p_simt = nb.PW_NN.batch_eval(model, None, big, pool[:40], ps, 64, stats, 'posteriors')[0]
    finally:
        eng.set_tensor_cores(True)
    p_ora = O.batch_eval(layers, w, big, pool[:40], ps, 64, stats, 'posteriors')[0]
    assert np.abs(p_simt - p_ora).max() < 1e-4


def test_overflow_flag_isolated_layers(nb):
    eng = nb.get_engine()
    from nnal_b200 import _lib as L
    rs = np.random.RandomState(0)
    A = rs.randn(256, 128).astype(np.float32)
    W = rs.randn(64, 128).astype(np.float32)
    b = np.zeros(64, np.float32)
    eng.debug_fc(A, W, b, 1, 1)
    A[3, 5] = 1e6
    with pytest.raises(L.NnalOverflowError):
        eng.debug_fc(A, W, b, 1, 1)
    A[3, 5] = np.nan
    with pytest.raises(L.NnalOverflowError):
        eng.debug_fc(A, W, b, 1, 1)
    out = eng.debug_fc(A, W, b, 0, 0)                         # FP32 CUDA-core GEMM (no activation): NaN propagates, no error
    assert np.isnan(out[3]).all() and np.isfinite(out[4]).all()


def test_volume_cache_sees_in_place_edits(nb):
    """The volume cache is keyed by a hash of the FULL content: an in-place edit of one voxel anywhere is uploaded."""
    ps, padded, stats, pool, layers, w = _setup(seed=9)
    model = nb.NN.create_PW1(2)
    model.set_weights(w)
    eng = nb.get_engine()
    eng.volume_cache = True
    expr = Expr(k=10, B=50, lambda_=0., patch_shape=ps, ntb=64, stats=stats)
    nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
    h0 = eng.h2d_bytes
    nb.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
    assert eng.h2d_bytes - h0 < sum(p.nbytes for p in padded)        # unchanged volumes: not uploaded again
    # edit ONE voxel inside the patch of pool sample 0 (far from any sampled checksum position)
    x, y, z = np.unravel_index(pool[0], (30, 28, 4))
    padded[1][x + 12, y + 12, z] += 50.
    g = nb.patch_utils.get_patches(padded, pool[:1], ps)
    assert np.array_equal(g, O.get_patches(padded, pool[:1], ps))
    p_dev = nb.PW_NN.batch_eval(model, None, padded, pool[:8], ps, 64, stats, 'posteriors')[0]
    p_ora = O.batch_eval(layers, w, padded, pool[:8], ps, 64, stats, 'posteriors')[0]
    assert np.abs(p_dev - p_ora).max() < 1e-4
    # equal content in NEW array objects is a hit
    copies = [p.copy() for p in padded]
    h0 = eng.h2d_bytes
    nb.patch_utils.get_patches(copies, pool[:1], ps)
    assert eng.h2d_bytes - h0 < padded[0].nbytes


def test_adapter_weights_are_reread_every_query(nb):
    """A live reference model fine-tuned between two queries (its variables change, nobody calls refresh()) is scored
    with the NEW weights."""
    from collections import OrderedDict
    ps, padded, stats, pool, layers, w = _setup(seed=21)

    class Var(object):
        def __init__(self, a):
            self.a = a

        def eval(self, session=None):
            return self.a

    class RefModel(object):
        pass
    ref = RefModel()
    ref.var_dict = {name: [Var(np.array(W)), Var(np.array(b))] for name, (W, b) in w.items()}
    ref.dropout_rate = 1.
    adapter = nb.NN.ReferenceModelAdapter(ref, OrderedDict(layers), (25, 25, 3), feature_layer=len(layers) - 2)
    p1 = nb.PW_NN.batch_eval(adapter, None, padded, pool[:20], ps, 64, stats, 'posteriors')[0]
    assert np.abs(p1 - O.batch_eval(layers, w, padded, pool[:20], ps, 64, stats, 'posteriors')[0]).max() < 1e-4
    # "fine-tune": the variables' content changes in place
    w2 = {k: (v[0].copy(), v[1].copy()) for k, v in w.items()}
    w2['fc3'] = (w2['fc3'][0] * np.float32(0.5), w2['fc3'][1] + np.float32(0.3))
    ref.var_dict['fc3'][0].a[...] = w2['fc3'][0]
    ref.var_dict['fc3'][1].a[...] = w2['fc3'][1]
    p2 = nb.PW_NN.batch_eval(adapter, None, padded, pool[:20], ps, 64, stats, 'posteriors')[0]
    assert np.abs(p2 - O.batch_eval(layers, w2, padded, pool[:20], ps, 64, stats, 'posteriors')[0]).max() < 1e-4
    assert np.abs(p2 - p1).max() > 1e-3


def test_device_topk_merge_matches_global_stable_sort(nb):
    """pool_topk_device on several contexts + merge of the gathered pair buffers == np.argsort(kind='stable') of the
    concatenated scores (ties across ranks -> lowest global position), padding slots dropped."""
    import torch
    rs = np.random.RandomState(5)
    n_per = [700, 0, 1300, 90]
    scores = [np.round(rs.rand(n), 2) for n in n_per]                 # many exact ties
    scores[2][:5] = scores[0][:5]
    lo = np.concatenate([[0], np.cumsum(n_per)])
    allsc = np.concatenate(scores)
    for k in (1, 64, 100, 500):
        bufs = []
        engs = [nb.Engine(0) for _ in n_per]
        try:
            for r, e in enumerate(engs):
                buf = torch.zeros(k * 16, dtype=torch.uint8, device='cuda')
                # a pool pass is emulated by loading the scores: model with one fc layer is overkill -> use the score hook
                e._load_scores_for_test(scores[r])
                e.pool_topk_device(k, k, int(lo[r]), buf.data_ptr())
                e.synchronize()
                bufs.append(buf)
            recv = torch.cat(bufs)
            torch.cuda.synchronize()
            pos, sc = engs[0].topk_merge_pairs(recv.data_ptr(), len(n_per) * k, k)
        finally:
            for e in engs:
                e.close()
        want = np.argsort(allsc, kind='stable')[:k]
        assert np.array_equal(pos, want)
        assert np.array_equal(sc, allsc[want])
    # fewer valid entries than k: padding is dropped
    e = nb.Engine(0)
    try:
        buf = torch.zeros(50 * 16, dtype=torch.uint8, device='cuda')
        e._load_scores_for_test(scores[3])
        e.pool_topk_device(50, 50, 7, buf.data_ptr())
        e.synchronize()                                   # the library's stream is not torch's: order the two by hand
        two = torch.cat([buf, buf.clone()])
        two[50 * 16:].view(torch.float64).view(-1, 2)[:, 0] = float('inf')
        two[50 * 16:].view(torch.int64).view(-1, 2)[:, 1] = np.iinfo(np.int64).max
        torch.cuda.synchronize()
        pos, sc = e.topk_merge_pairs(two.data_ptr(), 100, 100)
    finally:
        e.close()
    assert len(pos) == 50 and np.array_equal(pos, 7 + np.argsort(scores[3], kind='stable')[:50])


@pytest.mark.parametrize('dtype', [np.float32, np.float64])
def test_large_pageable_volume_upload_is_exact(nb, dtype):
    """Volumes beyond 16 MB in pageable host memory go through the threaded pinned staging (4 MB pieces on 8 copy
    streams); the gathered patches are bit-identical to NumPy slicing, and to the plain single-copy upload."""
    ps = (25, 25, 1)
    shape = (131, 127, 97) if dtype == np.float32 else (101, 97, 67)     # odd sizes: the last piece of a modality is partial
    rs = np.random.RandomState(5)
    imgs = [rs.standard_normal(shape).astype(dtype) for _ in range(3)]
    padded = pad_imgs(imgs, ps)
    assert sum(p.nbytes for p in padded) > (16 << 20)
    pool = rs.choice(int(np.prod(shape)), 300, replace=False).astype(np.int64)
    eng = nb.get_engine()
    eng.volume_cache = False
    try:
        got = nb.patch_utils.get_patches(padded, pool, ps)
        eng.debug_option('plain_upload', 1)
        plain = nb.patch_utils.get_patches(padded, pool, ps)
    finally:
        eng.debug_option('plain_upload', 0)
        eng.volume_cache = True
    want = O.get_patches(padded, pool, ps)
    assert got.dtype == want.dtype and np.array_equal(got, want) and np.array_equal(plain, want)


def test_peer_memory_allgather_protocol(nb):
    """csrc/p2p.cu, the all-gather of the greedy step messages over peer memory: three contexts on one GPU open each other's
    buffers (same-process form of the IPC set-up) and run 40 back-to-back exchanges, nothing but stream order and the flags
    between them.  Every rank must see every rank's message of THAT exchange (two slot parities, monotonic flag values)."""
    import torch
    from nnal_b200 import dist
    world, nbytes, steps = 3, 4096 + 64, 40
    engs = [nb.Engine(0) for _ in range(world)]
    try:
        for r, e in enumerate(engs):
            e.p2p_alloc(world, r, nbytes)
        bases = [e.p2p_base() for e in engs]
        for e in engs:
            e.p2p_open_local(bases)
        # payload of (rank r, exchange q): bytes (r * 31 + q * 7 + i) mod 251, all prepared before the first launch
        i = torch.arange(nbytes, dtype=torch.int64, device='cuda')
        send = [[((r * 31 + q * 7 + i) % 251).to(torch.uint8) for q in range(steps)] for r in range(world)]
        got = [torch.zeros(steps, world * nbytes, dtype=torch.uint8, device='cuda') for _ in range(world)]
        torch.cuda.synchronize()
        streams = [torch.cuda.ExternalStream(e.stream) for e in engs]
        for q in range(steps):
            for r, e in enumerate(engs):
                ptr = e.p2p_allgather(send[r][q].data_ptr(), nbytes, 1000 + q)      # (sequence numbers need not start at 0)
                with torch.cuda.stream(streams[r]):
                    got[r][q].copy_(dist.device_view(ptr, (world * nbytes,), '|u1'))
        for e in engs:
            e.synchronize()
        want = torch.stack([torch.cat([send[r][q] for r in range(world)]) for q in range(steps)])
        for r in range(world):
            assert torch.equal(got[r], want), 'rank %d' % r
        with pytest.raises(ValueError):
            engs[0].p2p_allgather(send[0][0].data_ptr(), nbytes + 16, 0)          # larger than the slot
    finally:
        for e in engs:
            e.close()
