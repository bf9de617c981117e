#!/bin/bash
# The round's closing call: whole GPU suite, smoke(), the driver's bench line, ncu launch list of the same command (eager), conv traffic capture.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$TAG.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log | cut -c1-300
timeout -s KILL 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 200 gpurun_out/bench_$TAG.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --train-steps 2 --no-train-graph --no-graph --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu rc=$?"
python tools/show_bench.py gpurun_out/bench_$TAG.json
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$TAG.json'))
t = d.get('train') or {}
print('value_mode', d.get('value_mode'), 'eager ms', d.get('ms_per_step_eager'))
print('train eager %s graphed %s note %s' % (t.get('ms_per_step_eager'), t.get('ms_per_step_graphed'), t.get('graph_note')))
print('e2e', d['e2e']['value'], 'kp', (d.get('e2e_from_keypoints') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
for r in (d.get('lbs_sweep') or {}).get('rows', []):
    print('lbs B=%d %.1f us  %.3f of HBM  %.2f M bodies/s' % (r['batch'], r['us_cold_l2'], r['frac'], r['bodies_per_s'] / 1e6))
PY
bash tools/gpu_call7.sh $TAG
