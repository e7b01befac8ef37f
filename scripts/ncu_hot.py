#!/usr/bin/env python
"""Top stall lines of one kernel from an ncu report's source page (SASS view).
  python scripts/ncu_hot.py report.ncu-rep <kernel regex> [N]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
blocks = out.split('"Kernel Name"')
for blk in blocks[1:2]:
    lines = blk.split('\n')
    print('kernel', lines[0][:150])
    rd = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
    hdr = rd[0]
    rows = [dict(zip(hdr, r)) for r in rd[1:] if len(r) == len(hdr)]
    tot = sum(int(r['# Samples'] or 0) for r in rows)
    stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    agg = {h: sum(int(r[h] or 0) for r in rows) for h in stall_cols}
    print('total samples', tot, ' by reason:', ', '.join('%s=%d' % (k[6:], v) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for i, r in enumerate(rows):
        r['_i'] = i
    top = sorted(rows, key=lambda r: -int(r['# Samples'] or 0))[:N]
    for r in sorted(top, key=lambda r: r['_i']):
        why = sorted(((int(r[h] or 0), h[6:]) for h in stall_cols), reverse=True)[:2]
        print('%5d %5.1f%%  %-70s %s' % (r['_i'], 100. * int(r['# Samples']) / max(tot, 1), r['Source'][:70],
                                         ' '.join('%s:%d' % (w, c) for c, w in why if c)))
