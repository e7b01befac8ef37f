#!/usr/bin/env python
"""Warp-stall samples of one kernel grouped by the code between synchronisation instructions (a crude role/region map):
  python scripts/ncu_regions.py report.ncu-rep <kernel regex>"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + pat],
                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
blk = out.split('"Kernel Name"')[1]
lines = blk.split('\n')
rd = list(csv.reader(io.StringIO('\n'.join(lines[1:]))))
hdr = rd[0]
rows = [dict(zip(hdr, r)) for r in rd[1:] if len(r) == len(hdr)]
tot = sum(int(r['# Samples'] or 0) for r in rows)
print('total samples', tot, 'instructions', sum(int(r['Instructions Executed'] or 0) for r in rows))
KEYS = ['SYNCS', 'UTCHMMA', 'BAR.', 'EXIT', 'NANOSLEEP', 'UTCBAR', 'LDTM', 'UBLKCP', 'UTMALDG', 'STG', 'LDG']
prev = 0
for i, r in enumerate(rows):
    s = r['Source']
    if any(k in s for k in KEYS):
        seg = sum(int(rows[j]['# Samples'] or 0) for j in range(prev, i))
        ex = sum(int(rows[j]['Instructions Executed'] or 0) for j in range(prev, i))
        if seg + int(r['# Samples'] or 0) > tot * 0.002:
            print('%5d  before: %6d samples %9d instr | self %5s  %s' % (i, seg, ex, r['# Samples'], s[:90]))
        prev = i + 1
