"""SASS evidence for profiles/: which kernels of the in-tree libdlsg.so contain tcgen05 (UTCHMMA), tensor-memory loads (LDTM),
TMA tensor loads (UTMALDG), bulk copies (UBLKCP), mbarrier ops (SYNCS), legacy mma (HMMA) / ldmatrix (LDSM) - static counts
from `cuobjdump -sass` (runs in the build container, no GPU).   python tools/sass_listing.py > profiles/rNN_sass_mnemonics.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'd-lsg-video-caption_b200', 'dlsg', 'libdlsg.so')
PAT = re.compile(r'\b(UTCHMMA|UTCQMMA|UTCBAR|UTMALDG|UTMASTG|UTMAPF|LDTM|STTM|UTCATOMSWS|UBLKCP|SYNCS|HMMA|LDSM|UTCCP|ACQBULK|REDG|UTMACCTL)[\.\w]*')
TWO = ('UTMALDG', 'HMMA', 'SYNCS', 'UBLKCP', 'LDSM')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            for mm in PAT.finditer(line):
                parts = mm.group(0).split('.')
                key = parts[0] + ('.' + parts[1] if len(parts) > 1 and parts[0] in TWO else '')
                cnt[cur][key] += 1
    print('# SASS mnemonics per kernel of libdlsg.so (cuobjdump -sass, sm_100a; static instruction counts).')
    print('# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, UBLKCP = cp.async.bulk,')
    print('# SYNCS = mbarrier, HMMA/LDSM = mma.sync / ldmatrix (the class (b) aggregation kernels), REDG = red.global')
    for f in sorted(cnt):
        c = cnt[f]
        if any(k.startswith(('UTC', 'UTMA', 'LDTM', 'UBLKCP', 'HMMA', 'LDSM')) for k in c):
            dem = subprocess.run(['c++filt', f], capture_output=True, text=True).stdout.strip()
            print('%s\n    %s' % (dem[:200], '  '.join('%s x%d' % (k, v) for k, v in sorted(c.items()))))
    return 0


if __name__ == '__main__':
    sys.exit(main())
