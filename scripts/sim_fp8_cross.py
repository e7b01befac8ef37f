"""CPU emulation (exact float64 accumulation) of operand formats for the tensor-core layers of PW1:
   f16x3   : x = hi + lo fp16 terms, products hi.hi + hi.lo + lo.hi          (shipped in round 1)
   f16+f8  : hi.hi in fp16, cross terms hi.lo + lo.hi with e4m3 operands      (kind::f8f6f4 at 2x the rate)
Reports max |posterior - float64 oracle| so the 1e-4 tolerance can be checked before writing kernels."""
import sys
import numpy as np
import torch
import torch.nn.functional as F
sys.path.insert(0, '.')
import oracle as O
from tests.util import centered_weights, pad_imgs, synth_volume, vol_stats

torch.set_num_threads(8)
E4 = torch.float8_e4m3fn


def f16(t):
    return t.to(torch.float16).double()


def f8(t):
    return t.clamp(-448, 448).float().to(E4).double()


def pow2_scale(W):
    m = float(W.abs().max())
    return 2.0 ** np.floor(np.log2(16384.0 / m))      # max |W*s| in [8192, 16384)


def lin(x, W, mode, conv=None):
    """x: activations (double), W: weights (double).  Returns the emulated product (double)."""
    op = (lambda a, w: F.conv2d(a, w, padding=w.shape[-1] // 2)) if conv else (lambda a, w: a @ w.t())
    if mode == 'f64':
        return op(x, W)
    s = pow2_scale(W)
    Ws = W * s
    xh = f16(x); xl = x - xh
    wh = f16(Ws); wl = Ws - wh
    if mode == 'f16x1':
        return op(xh, wh) / s
    if mode == 'f16x3':
        return (op(xh, wh) + op(xh, f16(wl)) + op(f16(xl), wh)) / s
    if mode.startswith('f8'):
        # cross terms in e4m3: [x_lo * 2^sa | x_hi] . [w_hi * 2^-sa ; w_lo]
        sa = float(mode.split(':')[1]) if ':' in mode else 10.
        c1 = op(f8(xl * 2.0 ** sa), f8(wh * 2.0 ** -sa))
        c2 = op(f8(xh), f8(wl))
        return (op(xh, wh) + c1 + c2) / s
    raise ValueError(mode)


def forward(layers, w, x, modes):
    """x [n,25,25,3] float64 numpy; modes: dict layer-name -> mode."""
    t = torch.from_numpy(x).double().permute(0, 3, 1, 2)        # NCHW
    for name, spec in layers:
        kind = spec[1]
        if kind == 'conv':
            W, b = w[name]
            Wt = torch.from_numpy(W).double().permute(3, 2, 0, 1)   # [cout][cin][kh][kw]
            t = lin(t, Wt, modes.get(name, 'f64'), conv=True) + torch.from_numpy(np.asarray(b).reshape(-1)).double().view(1, -1, 1, 1)
            t = t.clamp_min(0)
        elif kind == 'pool':
            t = F.max_pool2d(t, 2, 2, ceil_mode=True)
        elif kind == 'fc':
            W, b = w[name]
            if t.dim() == 4:
                # TF flatten: transpose reverses axes [N,H,W,C]->[C,W,H,N]; row = c*(W*H) + w*H + h
                t = t.permute(0, 1, 3, 2).reshape(t.shape[0], -1)
            Wt = torch.from_numpy(W).double()
            t = lin(t, Wt, modes.get(name, 'f64')) + torch.from_numpy(np.asarray(b).reshape(-1)).double().view(1, -1)
            if name != layers[-1][0]:
                t = t.clamp_min(0)
    return torch.softmax(t, 1).numpy()


def main():
    ps = (25, 25, 1)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
    imgs = synth_volume((48, 40, 6), 3, 0)
    padded = pad_imgs(imgs, ps)
    stats = vol_stats(imgs)
    pool = np.random.RandomState(1).choice(48 * 40 * 6, n, replace=False).astype(np.int64)
    layers = O.pw1_layers(2)
    x = O.normalize_batch_eval(O.get_patches(padded, pool, ps), stats)
    wts = centered_weights(layers, (25, 25, 3), 2, x[:64].astype(np.float32))
    print([(nm, sp) for nm, sp in layers])
    ref = O.forward(layers, wts, x)['posteriors'].T
    names = [nm for nm, sp in layers if sp[1] in ('conv', 'fc')]
    base = forward(layers, wts, x, {})
    print('torch f64 vs oracle f64: %.3g' % np.abs(base - ref).max())
    def run(tag, modes):
        p = forward(layers, wts, x, modes)
        print('%-44s max |dp| %.3g' % (tag, np.abs(p - base).max()))
    tc = names[:-1]                                    # the last fc (c outputs) runs on CUDA cores in fp32
    run('f16x1 all tc layers', {k: 'f16x1' for k in tc})
    run('f16x3 all tc layers', {k: 'f16x3' for k in tc})
    fcs = [k for k in tc if dict(layers)[k][1] == 'fc']
    for sa in (8., 10., 12.):
        m = {k: 'f16x3' for k in tc}
        m.update({k: 'f8:%g' % sa for k in fcs})
        run('f16x3 convs, f16+f8(e4m3, sa=%g) fcs' % sa, m)
    run('f16+f8 all tc layers (sa=10)', {k: 'f8:10' for k in tc})


if __name__ == '__main__':
    main()
