#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove the Blackwell-native path (profiles/rNN_sass_summary.md):
UTC*MMA (tcgen05.mma), LDTM/STTM (tcgen05.ld/st), UTMALDG/UTMASTG/UBLKCP (TMA), SYNCS (mbarrier), HMMA (legacy mma.sync).

  python scripts/sass_summary.py [path/to/libnnal_b200.so] > profiles/r2_sass_summary.md
"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'nn-active-learning_b200', 'libnnal_b200.so')
out = subprocess.run(['cuobjdump', '-sass', so], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
PAT = OrderedDict([('UTC*MMA', r'\bUTC[A-Z]*MMA\b'), ('LDTM', r'\bLDTM\b'), ('STTM', r'\bSTTM\b'), ('UTMALDG', r'\bUTMALDG\b'),
                   ('UTMASTG', r'\bUTMASTG\b'), ('UBLKCP', r'\bUBLKCP\b'), ('SYNCS', r'\bSYNCS\b'), ('HMMA', r'\bHMMA\b'),
                   ('DFMA', r'\bDFMA\b'), ('STG.128', r'\bSTG\.E\.128\b'), ('ATOMS', r'\bATOMS\b')])
kern = None
rows = OrderedDict()
total = {}
for line in out.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        name = subprocess.run(['c++filt', m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r'\(.*', '', name).replace('void ', '')
        kern = name
        rows.setdefault(kern, {k: 0 for k in PAT})
        total[kern] = 0
        continue
    if kern and re.match(r'\s+/\*[0-9a-f]{4,}\*/', line):
        total[kern] += 1
        for k, p in PAT.items():
            if re.search(p, line):
                rows[kern][k] += 1
print('# SASS summary of `libnnal_b200.so` (sm_100a) -- `python scripts/sass_summary.py`\n')
print('Counts of instructions per kernel (static code, `cuobjdump -sass`).  `UTC*MMA` = tcgen05.mma, `LDTM`/`STTM` = tcgen05.ld/st, '
      '`UTMALDG`/`UTMASTG`/`UBLKCP` = TMA tensor load / tensor store / bulk copy, `SYNCS` = mbarrier, `HMMA` = legacy mma.sync, '
      '`DFMA` = float64 FMA.\n')
print('| kernel | SASS instrs | ' + ' | '.join(PAT) + ' |')
print('|---|---|' + '---|' * len(PAT))
tot = {k: 0 for k in PAT}
for kname, r in rows.items():
    if not any(r.values()):
        continue
    print('| `%s` | %d | %s |' % (kname[:110], total[kname], ' | '.join(str(r[k]) if r[k] else '' for k in PAT)))
    for k in PAT:
        tot[k] += r[k]
print('| **all %d kernels** | %d | %s |' % (len(rows), sum(total.values()), ' | '.join(str(tot[k]) for k in PAT)))
