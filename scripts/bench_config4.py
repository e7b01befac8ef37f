#!/usr/bin/env python
"""Config 4 of BASELINE.json (memory-bound path): (i) gather the 25x25x3 patches of EVERY voxel of a
256x256x180 3-modality volume, slice by slice as PW_analyze_results.full_slice_eval does
(PW_analyze_results.py:689-715): 11.8 M patches, 88.5 GB of float32 patches if materialised -- here each
slice's 65,536 patches (491 MB) are written to a reused device buffer; (ii) pixel-wise entropy of a
[c=2,256,256,180] float32 posterior tensor.  CUDA-event timing on the library's stream; prints one JSON line.

  python scripts/bench_config4.py [--slices 180] [--reps 5]
"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench as Bn
import nnal_b200
from nnal_b200 import _lib as L

ap = argparse.ArgumentParser()
ap.add_argument('--slices', type=int, default=180)
ap.add_argument('--reps', type=int, default=5)
args = ap.parse_args()
peaks = Bn.load_peaks()
padded, stats, _ = Bn.make_workload(10)
eng = nnal_b200.get_engine()
eng.upload(0, padded)
st = np.array(stats, dtype=np.float64)
X, Y, Z = Bn.VOL_SHAPE
stream = torch.cuda.ExternalStream(eng.stream)
per_slice = X * Y
out = torch.empty((per_slice, 25, 25, 3), dtype=torch.float32, device='cuda')
# voxel ids of slice z: every (x, y) at fixed z (raveled C-order over (X,Y,Z))
xy = torch.arange(per_slice, dtype=torch.int64, device='cuda') * Z
torch.cuda.synchronize()

def gather_all(norm):
    for z in range(args.slices):
        inds = xy + z
        eng.gather_device(0, inds.data_ptr(), per_slice, Bn.PATCH, st, norm, out.data_ptr())

res = {}
for name, norm in (('raw', L.NORM_NONE), ('normalised', L.NORM_BATCH_EVAL)):
    idx = [xy + z for z in range(args.slices)]           # index tensors built outside the timed region
    torch.cuda.synchronize(); eng.synchronize()
    for z in range(min(3, args.slices)):
        eng.gather_device(0, idx[z].data_ptr(), per_slice, Bn.PATCH, st, norm, out.data_ptr())
    eng.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for z in range(args.slices):
        eng.gather_device(0, idx[z].data_ptr(), per_slice, Bn.PATCH, st, norm, out.data_ptr())
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1)
    npatch = per_slice * args.slices
    gbs = npatch * 15008.0 / (ms * 1e-3) / 1e9
    res['gather_' + name] = {'patches': npatch, 'ms': ms, 'patches_per_s': npatch / (ms * 1e-3), 'algorithmic_GBps': gbs,
                             'frac_of_hbm_peak': gbs / peaks['hbm_gbs'], 'bytes_per_patch': 15008}
# spot check of the last slice against the oracle
import oracle as O
chk = (xy[:64] + (args.slices - 1)).cpu().numpy()
ref = O.normalize_batch_eval(O.get_patches(padded, chk, Bn.PATCH), stats).astype(np.float32)
assert np.array_equal(out[:64].cpu().numpy(), ref), 'gather mismatch'

# (ii) pixel-wise entropy.  One launch lasts ~25 us, less than the host-side launch latency of a ctypes call, so
# the launches are queued back to back over NSET rotating tensor sets (NSET x 141 MB >> 126 MB L2: every launch
# reads posteriors that are no longer cached) and the whole train is timed with one event pair.
n = X * Y * Z
NSET = 4
g = torch.Generator(device='cuda'); g.manual_seed(5)
posts, Hs = [], []
for _ in range(NSET):
    logits = torch.randn((2, n), generator=g, device='cuda', dtype=torch.float32)
    posts.append(torch.softmax(logits, dim=0).contiguous())
    Hs.append(torch.empty(n, dtype=torch.float32, device='cuda'))
torch.cuda.synchronize()
for i in range(NSET):
    eng.entropy_device(posts[i].data_ptr(), 2, n, 1e-7, Hs[i].data_ptr())
eng.synchronize()
R = 8 * NSET
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
gate = torch.empty(1 << 28, dtype=torch.uint8, device='cuda')
tot = 0.0
for _ in range(args.reps):
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        gate.zero_()                                     # keeps the GPU busy while the launch train is queued
    e0.record(stream)
    for i in range(R):
        eng.entropy_device(posts[i % NSET].data_ptr(), 2, n, 1e-7, Hs[i % NSET].data_ptr())
    e1.record(stream); e1.synchronize()
    tot += e0.elapsed_time(e1) / R
ms = tot / args.reps
gbs = n * 12.0 / (ms * 1e-3) / 1e9
p64 = posts[0].double()
Href = -(p64 * torch.log(p64)).sum(0)
err = float(((Hs[0].double() - Href).abs() / Href.abs().clamp_min(1e-12)).max())
assert err < 1e-3, err
res['entropy_map'] = {'voxels': n, 'ms': ms, 'algorithmic_GBps': gbs, 'frac_of_hbm_peak': gbs / peaks['hbm_gbs'],
                      'bytes_per_voxel': 12, 'max_rel_err_vs_f64': err,
                      'l2': '%d rotating tensor sets of 141 MB (larger than L2), %d launches per timed train' % (NSET, R)}
res['peak_hbm_GBps'] = peaks['hbm_gbs']
res['peak_source'] = peaks['source']
print(json.dumps({'config4': res}))
