"""Multi-GPU plumbing: one process per GPU, pool positions block-partitioned across ranks
(SURVEY.md §8e), tiny collectives through ``torch.distributed`` (NCCL over NVLink on the GPU
box, gloo in CPU tests).  Scoring itself needs no data-path collective: every patch is scored
independently; only the selection couples ranks."""
import numpy as np


def _td():
    import torch.distributed as td
    return td


def is_dist():
    try:
        td = _td()
    except Exception:
        return False
    return td.is_available() and td.is_initialized() and td.get_world_size() > 1


def rank_world():
    if not is_dist():
        return 0, 1
    td = _td()
    return td.get_rank(), td.get_world_size()


def shard_bounds(n, world):
    """Contiguous block partition of pool POSITIONS: rank r owns [b[r], b[r+1])."""
    base, rem = divmod(int(n), int(world))
    sizes = [base + (1 if r < rem else 0) for r in range(world)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def _device():
    import torch
    td = _td()
    if td.get_backend() == 'nccl':
        import os
        return torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0')))
    return torch.device('cpu')


def allgather_topk(scores, positions, k):
    """Every rank contributes its local top-k (ascending scores, global positions); returns the
    global k smallest by (score, position) -- identical on every rank."""
    import torch
    if not is_dist():
        order = np.lexsort((positions, scores))[:k]
        return positions[order], scores[order]
    td = _td()
    world = td.get_world_size()
    dev = _device()
    s = np.full(k, np.inf, dtype=np.float64)
    p = np.full(k, np.iinfo(np.int64).max, dtype=np.int64)
    s[:len(scores)] = scores[:k]
    p[:len(positions)] = positions[:k]
    ts = torch.from_numpy(s).to(dev)
    tp = torch.from_numpy(p).to(dev)
    gs = [torch.empty_like(ts) for _ in range(world)]
    gp = [torch.empty_like(tp) for _ in range(world)]
    td.all_gather(gs, ts)
    td.all_gather(gp, tp)
    S = torch.cat(gs).cpu().numpy()
    P = torch.cat(gp).cpu().numpy()
    valid = P != np.iinfo(np.int64).max
    S, P = S[valid], P[valid]
    order = np.lexsort((P, S))[:k]
    return P[order], S[order]


def engine_stream(eng):
    """Context manager making the engine's CUDA stream torch's current stream, so that NCCL collectives issued
    inside it are ordered with the library's kernels (no host synchronisation)."""
    import contextlib
    import torch
    if getattr(eng, 'message_device', 'cpu') != 'cuda':
        return contextlib.nullcontext()
    return torch.cuda.stream(torch.cuda.ExternalStream(eng.stream))


class _DevPtr(object):
    """Exposes a raw device pointer of the library to torch (``torch.as_tensor(_DevPtr(...), device='cuda')``)
    through the CUDA array interface -- no copy; the engine keeps the memory alive."""

    def __init__(self, ptr, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {'shape': tuple(shape), 'typestr': typestr, 'data': (int(ptr), False),
                                         'version': 2, 'strides': strides}


def device_view(ptr, shape, typestr='<f4'):
    import torch
    return torch.as_tensor(_DevPtr(ptr, shape, typestr), device='cuda')


def topk_global(eng, k, lo, n_total):
    """The k pool samples with the smallest (score, global position) over ALL ranks' current pool scores
    (``np.argsort(score, kind='stable')[:k]`` of the concatenated pool, PW_NNAL.py:724-730); ``lo`` = global
    position of this rank's first sample.  Returns (positions, scores), identical on every rank.
    NCCL: every rank leaves its k best (score, position) pairs on the device, ONE packed
    ``all_gather_into_tensor`` on the engine's stream, device-side merge, one k-element read-back."""
    k = int(min(max(int(k), 0), int(n_total)))
    if not is_dist():
        idx, sc = eng.pool_topk(k, with_scores=True)
        return idx + lo, sc
    td = _td()
    if td.get_backend() != 'nccl' or not hasattr(eng, 'pool_topk_device'):
        idx, sc = eng.pool_topk(k, with_scores=True)          # gloo (CPU tests over the NumPy fake engine)
        return allgather_topk(sc, idx + lo, k)
    import torch
    world = td.get_world_size()
    if k == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0)
    with engine_stream(eng):
        send = torch.empty(k * 16, dtype=torch.uint8, device='cuda')
        recv = torch.empty(world * k * 16, dtype=torch.uint8, device='cuda')
        eng.pool_topk_device(k, k, lo, send.data_ptr())
        td.all_gather_into_tensor(recv, send)
        pos, sc = eng.topk_merge_pairs(recv.data_ptr(), world * k, k)
    return pos, sc


def allreduce_device_f32_(eng, ptr, count):
    """In-place NCCL sum all-reduce of ``count`` float32 values at device pointer ``ptr`` (library memory), on
    the engine's stream.  Returns the bytes reduced (0 in a single process)."""
    if not is_dist():
        return 0
    with engine_stream(eng):
        t = device_view(ptr, (int(count),), '<f4')
        _td().all_reduce(t)
    return int(count) * 4


def allgather_concat(arr):
    """Concatenate per-rank 1-D arrays (rank order) on every rank."""
    import torch
    if not is_dist():
        return arr
    td = _td()
    world = td.get_world_size()
    dev = _device()
    n = torch.tensor([len(arr)], dtype=torch.int64, device=dev)
    ns = [torch.empty_like(n) for _ in range(world)]
    td.all_gather(ns, n)
    ns = [int(x.item()) for x in ns]
    mx = max(ns) if ns else 0
    buf = np.zeros(mx, dtype=arr.dtype)
    buf[:len(arr)] = arr
    t = torch.from_numpy(buf).to(dev)
    g = [torch.empty_like(t) for _ in range(world)]
    td.all_gather(g, t)
    return np.concatenate([x.cpu().numpy()[:m] for x, m in zip(g, ns)])


def allreduce_argmin(value, payload):
    """Global (min value, lowest payload on ties) over ranks; returns (value, payload, owner)."""
    import torch
    if not is_dist():
        return value, payload, 0
    td = _td()
    world = td.get_world_size()
    dev = _device()
    t = torch.tensor([float(value), float(payload)], dtype=torch.float64, device=dev)
    g = [torch.empty_like(t) for _ in range(world)]
    td.all_gather(g, t)
    vals = np.array([[x[0].item(), x[1].item()] for x in g])
    order = np.lexsort((vals[:, 1], vals[:, 0]))
    o = int(order[0])
    return vals[o, 0], int(vals[o, 1]), o


def broadcast_array(arr, src):
    import torch
    if not is_dist():
        return arr
    td = _td()
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(_device())
    td.broadcast(t, src)
    return t.cpu().numpy()


def allreduce_sum_(tensor):
    """In-place sum all-reduce of a torch tensor (e.g. the per-GPU Gram partials)."""
    if is_dist():
        _td().all_reduce(tensor)
    return tensor
