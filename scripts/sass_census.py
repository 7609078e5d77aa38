#!/usr/bin/env python
"""Per-kernel SASS instruction census of the built library (cuobjdump -sass): total instruction count and the
Blackwell-specific mnemonics.  Usage: python scripts/sass_census.py > profiles/r2_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'tensorrec_b200', 'libtensorrec_b200.so')
WATCH = ('UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'UTCMMA', 'HMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS',
         'REDG', 'ATOMG', 'FMNMX3', 'FMNMX', 'SHFL', 'VOTE', 'MATCH', 'ELECT', 'UTCBAR')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], check=True, capture_output=True, text=True).stdout
    names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), check=True,
                           capture_output=True, text=True).stdout.split('\n')
    kernels, cur, i = [], None, 0
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = [re.sub(r'\(.*', '', names[i]), 0, collections.Counter()]
            kernels.append(cur)
            i += 1
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur is not None:
            cur[1] += 1
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    cur[2][w] += 1
                    break
    print('SASS instruction census of tensorrec_b200/libtensorrec_b200.so (cuobjdump -sass, sm_100a), per kernel: total '
          'instructions and the\nBlackwell-specific mnemonics (B200_PROFILING.md: tcgen05.mma = UTC*MMA, tcgen05.ld/st = '
          'LDTM/STTM, TMA = UTMALDG/UTMASTG/UBLKCP,\nmbarrier = SYNCS; no HMMA = no legacy mma.sync path).  Regenerate: '
          'python scripts/sass_census.py > profiles/r2_sass_census.txt\n')
    print('%-72s %7s  %s' % ('kernel', 'instrs', '(counts of the mnemonics that occur)'))
    for name, n, c in kernels:
        print('%-72s %7d  %s' % (name[:72], n, '  '.join('%s=%d' % (w, c[w]) for w in WATCH if c[w])))
    total = collections.Counter()
    for _, _, c in kernels:
        total.update(c)
    print('\n%d kernels; totals: %s' % (len(kernels), '  '.join('%s=%d' % (w, total[w]) for w in WATCH if total[w])))
    if total['HMMA']:
        sys.exit('HMMA found: a legacy mma.sync path is in the library')


if __name__ == '__main__':
    main()
