"""Decision table for the 2-MMA FC variant: max |posterior - float64| of the emulated operand formats
over seeds x weight sets, exact float64 accumulation (the tensor core's own accumulation error, ~1e-5 measured in
profiles/r1_precision.md, comes on top).  Writes a markdown table to stdout.

  python scripts/sim_fp8_table.py [n_patches]
"""
import sys
import numpy as np
sys.path.insert(0, '.')
sys.path.insert(0, 'scripts')
import oracle as O
import sim_fp8_cross as S
from tests.util import centered_weights, pad_imgs, synth_volume, vol_stats


def case_config2(seed, n):
    """PW1 c=2 on 25x25x3 patches of synthetic volumes (config 2): He-normal weights (bench seed family), centred head."""
    ps = (25, 25, 1)
    imgs = synth_volume((48, 40, 6), 3, seed)
    padded = pad_imgs(imgs, ps)
    stats = vol_stats(imgs)
    pool = np.random.RandomState(seed + 1).choice(48 * 40 * 6, n, replace=False).astype(np.int64)
    layers = O.pw1_layers(2)
    x = O.normalize_batch_eval(O.get_patches(padded, pool, ps), stats)
    w = centered_weights(layers, (25, 25, 3), seed + 2, x[:64].astype(np.float32))
    return layers, w, x


def case_config2_raw(seed, n):
    """Same pool, plain He-normal weights with zero biases exactly as bench.py / NN.initialize draws them."""
    layers, _, x = case_config2(seed, n)
    return layers, O.he_init_weights(layers, (25, 25, 3), 4 + seed), x


def case_config1(seed, n):
    """PW1 layer dict on 28x28x1 U[0,1) images, c = 10 (config 1)."""
    layers = O.pw1_layers(10)
    x = np.random.RandomState(seed).rand(n, 28, 28, 1)
    w = O.he_init_weights(layers, (28, 28, 1), seed + 1)
    return layers, w, x


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 800
    rows = []
    for cname, fn in (('config2 (centred head)', case_config2), ('config2 (bench init)', case_config2_raw), ('config1 (28x28x1, c=10)', case_config1)):
        for seed in (0, 10, 20):
            layers, w, x = fn(seed, n)
            names = [nm for nm, sp in layers if sp[1] in ('conv', 'fc')]
            tc = names[:-1]
            fcs = [k for k in tc if dict(layers)[k][1] == 'fc']
            base = S.forward(layers, w, x, {})
            res = {}
            res['f16x3'] = np.abs(S.forward(layers, w, x, {k: 'f16x3' for k in tc}) - base).max()
            for sa in (8., 10.):
                m = {k: 'f16x3' for k in tc}
                m.update({k: 'f8:%g' % sa for k in fcs})
                res['fc f8 sa=%g' % sa] = np.abs(S.forward(layers, w, x, m) - base).max()
            m = {k: 'f16x3' for k in tc}
            m.update({fcs[0]: 'f8:10'})
            res['fc1 only f8'] = np.abs(S.forward(layers, w, x, m) - base).max()
            rows.append((cname, seed, res))
            print('| %s | %d | %s |' % (cname, seed, ' | '.join('%.2g' % res[k] for k in res)), flush=True)
    print()
    keys = list(rows[0][2].keys())
    print('| weights / pool | seed | ' + ' | '.join(keys) + ' |')
    print('|---|---|' + '---|' * len(keys))
    for cname, seed, res in rows:
        print('| %s | %d | %s |' % (cname, seed, ' | '.join('%.2g' % res[k] for k in keys)))
    print('| **worst case** | | ' + ' | '.join('**%.2g**' % max(r[2][k] for r in rows) for k in keys) + ' |')


if __name__ == '__main__':
    main()
