"""Counts the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md) per kernel of the built library.

    python tools/sass_summary.py [path/to/libstraps_b200.so]      -> one line per kernel that has any of them

UTCHMMA = tcgen05.mma (kind::f16), UTMALDG = cp.async.bulk.tensor (TMA tile load), UBLKCP = cp.async.bulk (1-D bulk copy),
LDTM = tcgen05.ld (TMEM -> registers), UTCBAR = tcgen05.commit, SYNCS = mbarrier operations, UTCATOMSWS / UTCALLOC-class = TMEM allocation.
"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, 'straps-3dhumanshapepose_b200', 'straps_b200', 'libstraps_b200.so')
MNEMONICS = ('UTCHMMA', 'UTMALDG', 'UBLKCP', 'LDTM', 'UTCBAR', 'SYNCS', 'UTCATOMSWS', 'ACQBULK')


def summary(lib=LIB):
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and m.group(1) in MNEMONICS:
            counts[cur][m.group(1)] += 1
    return counts


def demangle(names):
    try:
        out = subprocess.run(['cu++filt'] + list(names), capture_output=True, text=True, check=True).stdout.splitlines()
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else LIB
    counts = {k: v for k, v in summary(lib).items() if v}
    names = demangle(list(counts))
    print('%-86s %s' % ('kernel', ' '.join('%8s' % m for m in MNEMONICS)))
    for k, c in counts.items():
        short = re.sub(r'\(.*', '', re.sub(r'\((int|bool)\)', '', names[k])).replace('void ', '')
        print('%-86s %s' % (short[:86], ' '.join('%8d' % c[m] for m in MNEMONICS)))


if __name__ == '__main__':
    main()
