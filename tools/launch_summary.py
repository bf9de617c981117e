"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list (tools/gpu_call.sh): per-kernel totals of one step.

    python tools/launch_summary.py gpurun_out/launches_TAG.csv [--from KERNEL_SUBSTR] [--to KERNEL_SUBSTR] [--nth N]
Takes the launches from the N-th last occurrence of --from (default: the last pack_input / pack_proxy kernel = start of the last
inference step) up to (not including) the next occurrence of --to (default: end of file) and prints them grouped by kernel name."""
import argparse
import collections
import csv
import re


def load(fn):
    with open(fn) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for r in csv.DictReader(lines):
        v = float(r['Metric Value'].replace(',', ''))
        u = r['Metric Unit']
        v = v / 1000 if u in ('nsecond', 'ns') else v * 1000 if u in ('msecond', 'ms') else v
        rows.append((r['Kernel Name'], v))
    return rows


def short(name):
    name = re.sub(r'\(.*', '', name).replace('void ', '')
    name = re.sub(r'at::native::', 'native::', name)
    return name[:90]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--from', dest='start', default='pack_')
    ap.add_argument('--to', dest='stop', default=None)
    ap.add_argument('--nth', type=int, default=1, help='N-th last occurrence of --from')
    ap.add_argument('--list', action='store_true', help='also print the launches in order')
    a = ap.parse_args()
    rows = load(a.csv)
    idx = [i for i, (n, _) in enumerate(rows) if a.start in n]
    s = idx[-a.nth]
    e = len(rows)
    if a.stop:
        later = [i for i, (n, _) in enumerate(rows) if i > s and a.stop in n]
        if later:
            e = later[0]
    seg = rows[s:e]
    tot = collections.OrderedDict()
    for n, v in seg:
        k = short(n)
        t = tot.setdefault(k, [0.0, 0])
        t[0] += v
        t[1] += 1
    total = sum(v for _, v in seg)
    print('launches %d   kernel time %.1f us' % (len(seg), total))
    for k, (v, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
        print('%9.1f us %5.1f%% x%3d  %s' % (v, 100 * v / total, c, k))
    if a.list:
        for n, v in seg:
            print('%9.1f  %s' % (v, short(n)))


if __name__ == '__main__':
    main()
