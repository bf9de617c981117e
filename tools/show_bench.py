"""Print the interesting fields of bench.py JSON lines: python tools/show_bench.py file.json [...]"""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get('roofline') or {}
        print('%-32s ms/step %.4f  encoder_ms %.4f  value %.0f  e2e %.0f  frac %.4f' % (
            f, d['ms_per_step'], r.get('encoder_ms', float('nan')), d['value'], (d.get('e2e') or {}).get('value', float('nan')),
            r.get('frac', float('nan'))))
    except Exception as e:   # noqa: BLE001 -- a failed run leaves an empty file; say so and go on
        print('%-32s unreadable (%s)' % (f, e))
