"""Summarise `ncu -i X.ncu-rep --page raw --csv` (one row per profiled launch): duration, DRAM bytes, pipe / cache utilisation.

    python tools/ncu_summary.py raw.csv [--match REGEX] [--traffic-json OUT.json --source TEXT]
--traffic-json writes the per-launch DRAM traffic of the matching launches (bench.py reports it as roofline.traffic)."""
import argparse
import csv
import json
import re

COLS = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'MB_rd'), ('dram__bytes_write.sum', 'MB_wr'),
        ('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', 'tensor%'),
        ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2%'), ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1%'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram%'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm%'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid')]
SCALE = {'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3, 'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('csv')
    ap.add_argument('--match', default='.')
    ap.add_argument('--traffic-json')
    ap.add_argument('--source', default='')
    a = ap.parse_args()
    rows = list(csv.reader(l for l in open(a.csv) if not l.startswith('==')))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ik = hdr.index('Kernel Name')
    idx = {}
    for name, _ in COLS:
        idx[name] = hdr.index(name) if name in hdr else None
    print('%-46s ' % 'kernel' + ' '.join('%8s' % s for _, s in COLS))
    tot = {'us': 0.0, 'MB_rd': 0.0, 'MB_wr': 0.0, 'n': 0}
    for r in data:
        if not re.search(a.match, r[ik]):
            continue
        vals = []
        for name, short in COLS:
            i = idx[name]
            if i is None or r[i] in ('', 'n/a'):
                vals.append(float('nan'))
                continue
            v = float(r[i].replace(',', '')) * SCALE.get(units[i], 1.0)
            vals.append(v)
            if short in tot:
                tot[short] += v
        tot['n'] += 1
        print('%-46s ' % re.sub(r'\(.*', '', r[ik]).replace('straps::', '')[:46] + ' '.join('%8.1f' % v for v in vals))
    print('launches %d  time %.1f us  DRAM read %.1f MB  write %.1f MB' % (tot['n'], tot['us'], tot['MB_rd'], tot['MB_wr']))
    if a.traffic_json and tot['n']:
        json.dump({'source': a.source, 'launches': tot['n'], 'dram_bytes_read_sum': tot['MB_rd'] * 1e6, 'dram_bytes_write_sum': tot['MB_wr'] * 1e6,
                   'dram_bytes_per_launch': (tot['MB_rd'] + tot['MB_wr']) * 1e6 / tot['n'], 'time_us_sum_under_ncu': tot['us']},
                  open(a.traffic_json, 'w'), indent=1)


if __name__ == '__main__':
    main()
