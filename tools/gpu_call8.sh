#!/bin/bash
# training-path tests, smoke(), short bench (training arm), then the conv-traffic capture
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py tests/test_gpu_checkpoint.py -m gpu -q > gpurun_out/pytest_train_$TAG.log 2>&1; echo "train tests rc=$?"; tail -4 gpurun_out/pytest_train_$TAG.log
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
timeout -s KILL 400 python bench.py --steps 10 --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$TAG.err
python tools/show_bench.py gpurun_out/bench_$TAG.json
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$TAG.json'))
t = d.get('train') or {}
print('train eager %s graphed %s note %s launches %s' % (t.get('ms_per_step_eager'), t.get('ms_per_step_graphed'), t.get('graph_note'), t.get('library_launches_per_step')))
PY
bash tools/gpu_call7.sh $TAG
