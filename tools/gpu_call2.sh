#!/bin/bash
# One gpurun call, targeted: the tests named first (fail fast), then the whole GPU suite, the bench line, the eager ncu launch list.
#   gpurun --timeout 1500 -- 'bash tools/gpu_call2.sh TAG "tests/test_gpu_smpl.py tests/test_gpu_train_step.py"'
TAG=${1:-r02}
FIRST=${2:-}
mkdir -p gpurun_out
if [ -n "$FIRST" ]; then
  timeout -s KILL 600 python -m pytest $FIRST -m gpu -q > gpurun_out/pytest_first_$TAG.log 2>&1; echo "first rc=$?"; tail -15 gpurun_out/pytest_first_$TAG.log
fi
timeout -s KILL 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_$TAG.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$TAG.err
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --train-steps 2 --no-train-graph --no-graph --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu rc=$?"
python tools/show_bench.py gpurun_out/bench_$TAG.json
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$TAG.json'))
t = d.get('train') or {}
print('train eager %s graphed %s note %s' % (t.get('ms_per_step_eager'), t.get('ms_per_step_graphed'), t.get('graph_note')))
for r in (d.get('lbs_sweep') or {}).get('rows', []):
    print('lbs B=%d %.1f us  %.3f of HBM  %.2f M bodies/s' % (r['batch'], r['us_cold_l2'], r['frac'], r['bodies_per_s'] / 1e6))
PY
