#!/bin/bash
# LBS-only call: parity tests, the sweep by mode, per-kernel durations (ncu launch list) at B = 64 and 4096.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_smpl.py -m gpu -q -x > gpurun_out/pytest_smpl_$TAG.log 2>&1; echo "smpl rc=$?"; tail -4 gpurun_out/pytest_smpl_$TAG.log
timeout -s KILL 300 python tools/bench_lbs.py --batches 1 8 32 64 128 --modes tc simt --iters 20 > gpurun_out/lbs_modes_$TAG.jsonl 2> gpurun_out/lbs_modes_$TAG.err; echo "lbs rc=$?"
timeout -s KILL 300 python tools/bench_lbs.py --batches 256 1024 4096 --modes tc --iters 20 >> gpurun_out/lbs_modes_$TAG.jsonl 2>> gpurun_out/lbs_modes_$TAG.err; echo "lbs rc=$?"
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/lbs_launches_$TAG.csv \
  python tools/bench_lbs.py --batches 64 4096 --modes tc --iters 2 > /dev/null 2>&1; echo "ncu rc=$?"
python - <<PY
import json, csv, collections
for l in open('gpurun_out/lbs_modes_$TAG.jsonl'):
    r = json.loads(l)
    print('mode %-5s B=%-4d cold %.1f us  graph cold %.1f us  graph b2b %.1f us  frac(graph cold) %.3f' % (r['lbs_mode'], r['batch'], 1e3 * r['ms_cold_l2'],
          1e3 * r.get('ms_graph_cold_l2', float('nan')), 1e3 * r.get('ms_graph_back_to_back', float('nan')), r.get('frac_graph_cold', float('nan'))))
rows = [r for r in csv.DictReader(l for l in open('gpurun_out/lbs_launches_$TAG.csv') if not l.startswith('=='))]
agg = collections.OrderedDict()
for r in rows:
    n = r['Kernel Name'][:40]
    if 'lbs' in n or 'chain' in n or 'joints' in n:
        agg.setdefault(n, []).append(float(r['Metric Value'].replace(',', '')))
for n, v in agg.items():
    print(n, ' '.join('%.1f' % (x / 1000 if x > 1000 else x) for x in v[-12:]))
PY
