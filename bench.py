#!/usr/bin/env python
"""bench.py -- headline benchmark of the query-scoring hot path (BASELINE.json metric: pool samples
scored per second for an entropy query round; config 2 of BASELINE.json at N=1).

A "step" is one query round over one synthetic pool: gather 25x25x3 patches around `--pool` voxels of
3 synthetic 256x256x180 volumes -> PW1 patch-CNN forward (c=2) -> |P(class1)-0.5| -> top-k (k=100).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  `value` = whole-job samples/s with volumes, weights and pool indices
resident in HBM; `e2e` = the same round through the reference-facing API
(nnal_b200.PW_NNAL.CNN_query with HOST arrays: volumes + indices copied host->device from pinned
memory and the selected indices read back, every step).  `--impl reference` times the restated
reference CPU path (oracle port: NumPy gather + float32 torch-CPU forward + NumPy scoring) on the
host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

VOL_SHAPE = (256, 256, 180)
PATCH = (25, 25, 1)
N_MOD = 3
METRIC = 'pool samples scored/sec (entropy query round: gather + PW1 forward + |p-0.5| + top-k)'
UNIT = 'samples/s'
PW1_MFLOP = 117.05          # per patch, SURVEY.md §8a row 5


class Expr(object):
    pass


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d.get('hbm_gbs', 6650.0), 'bf16_burst': d.get('bf16_tflops', 1590.0),
                'bf16_sustained': d.get('bf16_tflops_sustained', 1400.0), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_burst': 1590.0, 'bf16_sustained': 1400.0, 'source': 'fallback'}


def make_workload(pool_total, seed_pool=3, pinned=False):
    """config 2 of BASELINE.json / SURVEY.md §8d: m=3 volumes 256x256x180 float32 clip(N(100,30^2),0,inf),
    zero-padded by (12,12,0); stats = (mean,std) of the unpadded volume; pool = distinct raveled voxel ids."""
    g = np.random.Generator(np.random.PCG64(2))
    imgs, stats = [], []
    for j in range(N_MOD):
        v = g.standard_normal(VOL_SHAPE, dtype=np.float32)
        v *= 30.
        v += 100.
        np.maximum(v, 0, out=v)
        stats.append([float(v.mean(dtype=np.float64)), float(v.std(dtype=np.float64))])
        imgs.append(v)
    r = [(p - 1) // 2 for p in PATCH]
    padded = []
    for v in imgs:
        shp = tuple(v.shape[i] + 2 * r[i] for i in range(3))
        if pinned:
            import torch
            buf = torch.zeros(shp, dtype=torch.float32).pin_memory().numpy()
        else:
            buf = np.zeros(shp, dtype=np.float32)
        buf[r[0]:r[0] + v.shape[0], r[1]:r[1] + v.shape[1], r[2]:r[2] + v.shape[2]] = v
        padded.append(buf)
    nvox = int(np.prod(VOL_SHAPE))
    rs = np.random.RandomState(seed_pool)
    pool = rs.choice(nvox, pool_total, replace=False).astype(np.int64) if pool_total <= nvox else \
        rs.randint(0, nvox, pool_total).astype(np.int64)
    if pinned:
        import torch
        pb = torch.empty(pool_total, dtype=torch.int64).pin_memory().numpy()
        pb[:] = pool
        pool = pb
    return padded, stats, pool


def workload_config(args, world):
    return {'workload': 'config2: PW1 2-class CNN, 3x 256x256x180 f32 volumes, 25x25x3 patches, '
                        '%d-patch pool per GPU, entropy query k=100' % args.pool,
            'pool_per_gpu': args.pool, 'k': 100, 'parallelism': 'pool sharded x%d' % world,
            'l2': 'inputs larger than L2 (volumes 169 MB + >600 MB activations per chunk)',
            'forward_gflop_per_step_per_gpu': PW1_MFLOP * args.pool / 1e3}


def load_traffic():
    """DRAM bytes per sample of each kernel class, from the committed ncu --set full capture
    (profiles/r1_traffic.json, written by scripts/summarize_ncu.py traffic)."""
    p = os.path.join(ROOT, 'profiles', 'r1_traffic.json')
    return json.load(open(p)) if os.path.exists(p) else {}


def pw1_weights():
    import oracle as O
    layers = O.pw1_layers(2)
    return layers, O.he_init_weights(layers, (25, 25, 3), 4)


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks and throttle reasons DURING the timed region."""
    Q = 'index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,' \
        'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,' \
        'clocks_event_reasons.sw_power_cap'

    def __init__(self, gpu_index):
        threading.Thread.__init__(self)
        self.daemon = True
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[4:8]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_reference_round(sample, threads=None):
    """Restated reference CPU path on a bounded sample (oracle port; TF 1.x cannot be installed):
    NumPy gather (as patch_utils.get_patches) + float64 normalise + float32 torch-CPU forward + NumPy
    |p-0.5| argsort.  Returns (seconds, samples)."""
    import torch
    import oracle as O
    from oracle.torch_fp32 import TorchForward
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    padded, stats, pool = make_workload(sample)
    layers, w = pw1_weights()
    fwd = TorchForward(layers, w, feature_layer=len(layers) - 2, threads=threads)
    t0 = time.perf_counter()
    posts = O.batch_eval(layers, w, padded, pool, PATCH, 1000, stats, 'posteriors', fwd=fwd)[0]
    q = O.stable_topk(np.abs(posts - .5), 100)
    dt = time.perf_counter() - t0
    return dt, sample, threads, q


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = args.cpu_sample
    times = []
    threads = os.cpu_count()
    for i in range(args.warmup + args.steps):
        dt, n, threads, _ = cpu_reference_round(sample)
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * len(times) / total
    desc = '%d-patch sample of the %d-patch pool per step (NumPy gather + fp32 torch-CPU PW1 forward + NumPy ' \
           'argsort), restated reference CPU path (TF unavailable)' % (sample, args.pool)
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': workload_config(args, world),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': 'port', 'sample': desc},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))


def run_ours(args, rank, world):
    import torch
    import nnal_b200
    from nnal_b200 import _lib as L
    from nnal_b200 import dist

    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    peaks = load_peaks()
    pool_total = args.pool * world
    padded, stats, pool = make_workload(pool_total, pinned=True)
    layers, w = pw1_weights()
    model = nnal_b200.NN.create_PW1(2)
    model.set_weights(w)
    eng = nnal_b200.get_engine()
    eng.set_model(model)
    b = dist.shard_bounds(pool_total, world)
    lo, hi = int(b[rank]), int(b[rank + 1])
    n_local = hi - lo
    k = 100
    st = np.array(stats, dtype=np.float64)
    stream = torch.cuda.ExternalStream(eng.stream)
    d_inds = torch.from_numpy(np.ascontiguousarray(pool[lo:hi])).cuda()
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        eng.synchronize()
        torch.cuda.synchronize()

    # ---------------- device-resident leg (value) ----------------
    eng.upload(0, padded)

    def step_resident():
        eng.pool_begin(n_local, 0)
        eng.pool_eval_device(0, d_inds.data_ptr(), n_local, 0, PATCH, st)
        eng.pool_score(L.SCORE_BINARY)
        idx, sc = eng.pool_topk(k, with_scores=True)
        q, _ = dist.allgather_topk(sc, idx + lo, k)
        return q

    for _ in range(args.warmup):
        q_res = step_resident()
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    eng.profile(True)
    launches0 = eng.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        q_res = step_resident()
    ev1.record(stream)
    ev1.synchronize()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launches - launches0
    clocks = sampler.stop() if sampler else None
    # per kernel-class times measured over the same timed region
    classes = {}
    n_layers = len(layers)
    for i in range(n_layers):
        t, c = eng.profile_read(i)
        ty, macs, tc = eng.layer_info(i)
        classes[layers[i][0]] = {'ms': t, 'launches': c, 'macs_per_sample': macs, 'tc': tc, 'type': ty}
    for name, cid in (('gather', 100), ('score', 101), ('topk', 102)):
        t, c = eng.profile_read(cid)
        classes[name] = {'ms': t, 'launches': c, 'macs_per_sample': 0, 'tc': 0, 'type': -1}
    eng.profile(False)
    if world > 1:
        tms = torch.tensor([dev_ms], dtype=torch.float64, device='cuda')
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
        dev_ms = float(tms.item())
    value = pool_total * args.steps / (dev_ms * 1e-3)

    # ---------------- FI round (extra keys; the headline workload stays the entropy round) ----------------
    fi = None
    if args.fi_B > 0:
        fi = run_fi_round(args, eng, model, padded, stats, pool, lo, hi, d_inds, st, k, peaks, barrier, rank, world)

    mc = None
    if args.mc_T > 0:
        mc = run_mc_round(args, eng, model, lo, hi, d_inds, st, k)

    sdp = None
    if args.sdp_B > 0 and world == 1:
        sdp = run_sdp_round(args, eng, padded, pool, st, k)

    # ---------------- end-to-end leg through the reference-facing API ----------------
    expr = Expr()
    expr.pars = dict(k=k, B=k, lambda_=0., patch_shape=PATCH, ntb=10000, stats=stats)
    eng.volume_cache = False                      # volumes are copied host->device every step
    for _ in range(max(1, args.warmup // 2)):
        q_e2e = nnal_b200.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
    barrier()
    h0, d0 = eng.h2d_bytes, eng.d2h_bytes
    t0 = time.perf_counter()
    for _ in range(args.steps):
        q_e2e = nnal_b200.PW_NNAL.CNN_query(expr, model, None, padded, pool, None, 'entropy')
    eng.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    h2d = (eng.h2d_bytes - h0) / args.steps
    d2h = (eng.d2h_bytes - d0) / args.steps
    if world > 1:
        tms = torch.tensor([e2e_s], dtype=torch.float64, device='cuda')
        torch.distributed.all_reduce(tms, op=torch.distributed.ReduceOp.MAX)
        e2e_s = float(tms.item())
    e2e_value = pool_total * args.steps / e2e_s
    assert np.array_equal(np.sort(q_e2e), np.sort(q_res)), 'resident and e2e legs disagree'

    if rank != 0:
        return
    # ---------------- roofline of the dominant kernel class ----------------
    top = max(classes, key=lambda kk: classes[kk]['ms'])
    tinfo = classes[top]
    per_launch_ms = tinfo['ms'] / max(1, tinfo['launches'])
    samples_per_launch = n_local * args.steps / max(1, tinfo['launches'])
    if tinfo['macs_per_sample'] > 0:
        flops = 2.0 * tinfo['macs_per_sample'] * samples_per_launch
        achieved = flops / (per_launch_ms * 1e-3) / 1e12
        peak = peaks['bf16_sustained']
        tr = load_traffic().get(top)
        roofline = {'kernel': top, 'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                    'frac': achieved / peak,
                    'traffic': None if tr is None else tr['dram_bytes_per_sample'] * samples_per_launch,
                    'executed': None if not tinfo['tc'] else {'achieved': 3 * achieved, 'frac': 3 * achieved / peak,
                                                             'what': 'useful tensor FLOPs of the split-precision scheme: hi.hi + hi.lo + lo.hi per product'},
                    'note': 'algorithmic FLOPs (2*MACs) per launch / mean launch time; peak = %s sustained bf16; '
                            '%s' % (peaks['source'], 'tcgen05 fp16 hi/lo split executes 3x these FLOPs'
                                    if tinfo['tc'] else 'FP32 CUDA-core kernel (no tensor pipe)')}
    else:
        bytes_per = {'gather': 15008.0, 'score': 12.0, 'topk': 4.0}[top]
        achieved = bytes_per * samples_per_launch / (per_launch_ms * 1e-3) / 1e9
        tr = load_traffic().get(top)
        roofline = {'kernel': top, 'bound': 'hbm', 'achieved': achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                    'frac': achieved / peaks['hbm_gbs'],
                    'traffic': None if tr is None else tr['dram_bytes_per_sample'] * samples_per_launch,
                    'note': 'peak = %s copy bandwidth' % peaks['source']}
    stage_ms = {kk: round(v['ms'] / args.steps, 4) for kk, v in classes.items()}

    # ---------------- CPU baseline (bounded sample, rank 0, N=1 only) ----------------
    cpu = None
    if world == 1 and not args.no_cpu:
        dt, n, threads, _ = cpu_reference_round(args.cpu_sample)
        cpu = {'value': n / dt, 'unit': UNIT, 'cores': threads, 'kind': 'port',
               'sample': '%d-patch sample of the pool (NumPy gather + fp32 torch-CPU PW1 forward + NumPy argsort), '
                         'restated reference CPU path (TF 1.x unavailable)' % n}

    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f16x3 (fp16 hi/lo split operands, 3 tcgen05 MMAs per product, fp32 accumulate) + f32/f64 scoring',
            'data': 'synthetic',
            'config': workload_config(args, world),
            'clocks': clocks, 'gpu_launches': int(launches),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                    'ms_per_step': 1e3 * e2e_s / args.steps,
                    'api': 'nnal_b200.PW_NNAL.CNN_query(expr, model, sess, padded_imgs, pool_inds, tr_inds, "entropy")'},
            'roofline': roofline, 'stage_ms_per_step': stage_ms, 'cpu_baseline': cpu, 'fi_round': fi, 'mc_round': mc,
            'fi_sdp_round': sdp}
    print(json.dumps(line))


def run_mc_round(args, eng, model, lo, hi, d_inds, st, k):
    """One MC-entropy round (PW_NNAL.py:67-87) on the same pool: T dropout passes.  The conv trunk runs once per chunk,
    only the FC tail T times; the reference runs T complete batch_eval passes."""
    import torch
    from nnal_b200 import _lib as L
    n_local = hi - lo
    T = args.mc_T
    stream = torch.cuda.ExternalStream(eng.stream)
    eng.set_dropout_seed(5)

    def mc_step():
        eng.pool_mc_config(T, 0.5, model.dropout_layers, pos0=lo)
        try:
            eng.pool_begin(n_local, 0)
            eng.pool_eval_device(0, d_inds.data_ptr(), n_local, 0, PATCH, st)
        finally:
            eng.pool_mc_config(0, 1., [])
        eng.pool_score(L.SCORE_MC_BINARY)
        return eng.pool_topk(k)
    mc_step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    ev0.record(stream)
    for _ in range(reps):
        mc_step()
    ev1.record(stream)
    ev1.synchronize()
    ms = ev0.elapsed_time(ev1) / reps
    return {'method': 'MC-entropy', 'T': T, 'keep_prob': 0.5, 'ms_per_round': ms,
            'stochastic_passes_per_s': n_local * T / (ms * 1e-3),
            'note': 'conv trunk once per chunk + T FC-tail passes with Philox dropout fused into the FC epilogue / head'}


def run_sdp_round(args, eng, padded, pool, st, k):
    """The reference's literal FI selection (PW_NNAL.py:117-163) for B pre-filtered candidates: shrunk class-score
    gradients (one batched backward pass instead of 2B sess.run(tf.gradients) calls), A-matrices, SDP query distribution
    (first-order solver on the device, certified gap), sampling.  Host wall time per stage (the stages synchronise)."""
    from nnal_b200 import NNAL_tools
    B = min(args.sdp_B, len(pool))
    cand = pool[:B]
    res = None
    for rep in range(2):                       # first repetition warms the workspaces up
        eng.profile(True)
        t0 = time.perf_counter()
        post, g = eng.fi_shrunk_voxels(0, cand, PATCH, st, shape=padded[0].shape)
        t1 = time.perf_counter()
        fwd_ms, _ = eng.profile_read(120)
        bwd_ms, _ = eng.profile_read(121)
        eng.profile(False)
        t2 = time.perf_counter()
        r = eng.sdp_from_shrunk(g, post[1].astype(np.float64), 1e-5, tol=1e-4)       # A-matrices assembled on the device
        t3 = time.perf_counter()
        Q = NNAL_tools.sample_query_dstr(r['q'].copy(), k, replacement=True)
        t4 = time.perf_counter()
        res = {'B': int(B), 'k': int(k), 'tau': int(g.shape[2]), 'diag_load': 1e-5,
               'stage_ms': {'shrunk_gradients(gather+forward+backward)': 1e3 * (t1 - t0), 'sdp_solver(incl. A-matrices on the device)': 1e3 * (t3 - t2), 'sampling(host)': 1e3 * (t4 - t3)},
               'ms_per_round': 1e3 * (t4 - t0),
               'shrunk_device_ms': {'forward(all activations kept)': fwd_ms, 'backward(data gradients + layer sums)': bwd_ms},
               'backprops_per_s': 2.0 * B / (t1 - t0),
               'sdp': {'iterations': int(r['iterations']), 'objective': float(r['objective']), 'gap': float(r['gap']),
                       'us_per_iteration': 1e6 * (t3 - t2) / max(1, int(r['iterations'])),
                       'support': int((r['q'] > 1e-8).sum())},
               'n_selected': int(len(Q)),
               'note': 'reference: 2B single-sample tf.gradients runs over 36 M parameters + cvxopt SDP with an n x n '
                       'positivity block; here one batched data-gradient pass + tau x tau first-order solver'}
    if not args.no_cpu and res is not None:
        # the same selection through the float64 oracle on a bounded sample of the candidates (reported baseline: the
        # reference itself runs two TF sess.run(tf.gradients) calls per candidate, which cannot be installed here)
        import oracle as O
        ns = 48
        x = O.normalize_batch_eval(O.get_patches(padded, cand[:ns], PATCH), st).astype(np.float32)
        layers, w = pw1_weights()
        t0 = time.perf_counter()
        po, go = O.shrunk_class_gradients(layers, w, x)
        t1 = time.perf_counter()
        Ao = O.gen_A_matrices(go[0], go[1], po[1], 1e-5)
        qo, to, phio, gapo, ito = O.sdp_solve(Ao, 1e-4)
        t2 = time.perf_counter()
        res['cpu_baseline'] = {'kind': 'port', 'cores': 1, 'sample': '%d candidates (NumPy float64 forward + one backward pass per '
                               'class + shrink + multiplicative SDP)' % ns,
                               'backprops_per_s': 2.0 * ns / (t1 - t0), 'sdp_ms': 1e3 * (t2 - t1), 'sdp_iterations': int(ito)}
    return res


def run_fi_round(args, eng, model, padded, stats, pool, lo, hi, d_inds, st, k, peaks, barrier, rank, world):
    """One FI query round on the same pool (reference pipeline shape, PW_NNAL.py:89-163): pool pass ->
    uncertainty pre-filter to B -> second pass over the B candidates keeping the factors of the last two FC
    layers -> greedy k.  Also times the weighted Gram of the candidates (tensor cores).  Device timing per stage."""
    import torch
    import nnal_b200
    from nnal_b200 import _lib as L
    from nnal_b200 import dist, fi as fimod
    n_local = hi - lo
    B = min(args.fi_B, len(pool))
    delta = 1e-5
    stream = torch.cuda.ExternalStream(eng.stream)

    def fi_step():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record(stream)
        if B < len(pool):
            eng.pool_begin(n_local, 0)
            eng.pool_eval_device(0, d_inds.data_ptr(), n_local, 0, PATCH, st)
            eng.pool_score(L.SCORE_BINARY)
            idx, sc = eng.pool_topk(B, with_scores=True)
            sel, _ = dist.allgather_topk(sc, idx + lo, B)
        else:
            # B >= n: no pre-filter, every pool sample is a candidate (PW_NNAL.py:98-115 keeps all posteriors then)
            sel = np.arange(len(pool), dtype=np.int64)
        ev[1].record(stream)
        own = (sel >= lo) & (sel < hi)
        mine = sel[own]
        dbg = os.environ.get('NNAL_BENCH_DEBUG')
        t0 = time.perf_counter()
        eng.pool_begin(len(mine), 2)
        t1 = time.perf_counter()
        if len(mine):
            eng.pool_eval(0, pool[mine], 0, PATCH, st, shape=padded[0].shape)
        if dbg:
            eng.synchronize()
        t2 = time.perf_counter()
        eng.fi_set_candidates(None, 2)
        t3 = time.perf_counter()
        if dbg:
            sys.stderr.write('fi candidate pass: pool_begin %.2f ms, pool_eval %.2f ms, set_candidates %.2f ms\n'
                             % (1e3 * (t1 - t0), 1e3 * (t2 - t1), 1e3 * (t3 - t2)))
        ev[2].record(stream)
        chosen, obj = fimod.greedy_select(eng, k, delta, np.nonzero(own)[0].astype(np.int64))
        ev[3].record(stream)
        eng.fi_gram(None, read=False)
        ev[4].record(stream)
        ev[4].synchronize()
        return [ev[i].elapsed_time(ev[i + 1]) for i in range(4)], sel[chosen], obj, len(mine)

    for _ in range(max(1, args.warmup // 2)):
        fi_step()
    barrier()
    eng.profile(True)
    acc = np.zeros(4)
    steps = max(1, min(args.steps, 3))
    for _ in range(steps):
        t, q, obj, n_mine = fi_step()
        acc += np.array(t)
    gram_ms, gram_n = eng.profile_read(111)
    eng.profile(False)
    acc /= steps
    d = eng.fi_info()['d']
    gram_flops = 2.0 * n_mine * (d + 1) ** 2
    gram_ms_per = gram_ms / max(1, steps)
    if world > 1:
        tm = torch.tensor(list(acc), dtype=torch.float64, device='cuda')
        torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
        acc = tm.cpu().numpy()
    total = float(acc[:3].sum())
    return {'B': int(B), 'k': int(k), 'fi_layers': 2, 'delta': delta,
            'ms_per_round': total, 'samples_per_s': len(pool) / (total * 1e-3),
            'stage_ms': {'pool_pass+prefilter': float(acc[0]), 'candidate_pass+factors': float(acc[1]),
                         'greedy': float(acc[2]), 'gram(extra)': float(acc[3])},
            'greedy_us_per_step': 1e3 * float(acc[2]) / max(1, k),
            'gram': {'candidates_this_rank': int(n_mine), 'ms': gram_ms_per,
                     'tflops_algorithmic': gram_flops / max(gram_ms_per, 1e-9) / 1e9,
                     'frac_of_bf16_peak': gram_flops / max(gram_ms_per, 1e-9) / 1e9 / peaks['bf16_sustained'],
                     'note': 'full symmetric (d+1)^2 output, fp16 hi/lo split: 3 MMAs per product'},
            'objective_final': float(obj[-1]) if len(obj) else None}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--pool', type=int, default=100000, help='pool samples per GPU')
    ap.add_argument('--cpu-sample', type=int, default=5000)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--fi-B', type=int, default=10000, help='FI pre-filter size of the extra FI round (0: skip)')
    ap.add_argument('--sdp-B', type=int, default=10000, help='candidates of the extra literal-FI (shrunk gradients + SDP) round at 1 GPU (0: skip)')
    ap.add_argument('--mc-T', type=int, default=10, help='MC-dropout passes of the extra MC-entropy round (0: skip)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as td
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        td.init_process_group('nccl')
    try:
        run_ours(args, rank, world)
    finally:
        if world > 1:
            import torch.distributed as td
            td.destroy_process_group()


if __name__ == '__main__':
    main()
