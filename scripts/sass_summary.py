#!/usr/bin/env python
"""Static evidence for profiles/: per-kernel SASS mnemonic counts of the built library (cuobjdump -sass) and the ptxas
resource lines of the build logs.  Runs without a GPU:  python scripts/sass_summary.py > profiles/<tag>_sass_summary.md

What the mnemonics prove (B200 profiling recipe): UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor loads,
LDGSTS = cp.async, LDGMC = multimem.ld_reduce (NVSwitch in-fabric reduction), HMMA = the legacy mma.sync path (must be absent)."""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'mmd-gan_b200', 'libmmdgan_b200.so')
COLS = ['UTC*MMA', 'LDTM', 'UTMALDG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'LDGMC', 'HMMA', 'MUFU.EX2', 'SHFL', 'LDG.E.128', 'STG.E.128']


def demangle(names):
    out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if not m:
            continue
        op = m.group(1)
        c = counts[cur]
        c['total'] += 1
        if re.match(r'UTC[A-Z]*MMA', op):
            c['UTC*MMA'] += 1
        for key in ('LDTM', 'UTMALDG', 'UTCBAR', 'SYNCS', 'LDGSTS', 'LDGMC', 'HMMA', 'MUFU.EX2', 'SHFL', 'LDG.E.128', 'STG.E.128'):
            if op.startswith(key):
                c[key] += 1
    names = demangle(order)
    res = {}
    for log in sorted(glob.glob(os.path.join(ROOT, 'mmd-gan_b200', 'csrc', '*.ptxas.log'))):
        fn = None
        for line in open(log):
            m = re.search(r"Compiling entry function '(\S+)'", line)
            if m:
                fn = m.group(1)
            m = re.search(r'Used (\d+) registers', line)
            if m and fn:
                spill = re.search(r'(\d+) bytes spill stores', line)
                res[fn] = m.group(1)
    print('# Static SASS / ptxas summary of libmmdgan_b200.so (sm_100a; `python scripts/sass_summary.py`, no GPU needed)\n')
    print('`UTC*MMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, `UTCBAR` = tcgen05.commit, `SYNCS` = mbarrier,')
    print('`LDGSTS` = cp.async, `LDGMC` = multimem.ld_reduce (NVSwitch in-fabric reduction; multimem.st is an ordinary `STG...SYS` to the')
    print('multicast address), `HMMA` = legacy mma.sync (absent everywhere: no kernel falls back to the warp-level tensor path).\n')
    print('| kernel | regs | SASS instr | ' + ' | '.join(COLS) + ' |')
    print('|---|---|---|' + '---|' * len(COLS))
    tot = collections.Counter()
    for fn in order:
        c = counts[fn]
        tot.update(c)
        short = re.sub(r'\(.*', '', names.get(fn, fn)).replace('void ', '').replace('mg::', '')
        print('| `{}` | {} | {} | '.format(short, res.get(fn, ''), c['total']) + ' | '.join(str(c[k]) if c[k] else '' for k in COLS) + ' |')
    print('| **all {} kernels** | | {} | '.format(len(order), tot['total']) + ' | '.join(str(tot[k]) for k in COLS) + ' |')
    if tot['HMMA']:
        sys.exit('HMMA found: a kernel uses the legacy tensor path')


if __name__ == '__main__':
    main()
