#!/bin/bash
# One gpurun call: training-path tests, a short bench (training arm), and an `ncu --set full` capture of lbs_tc_kernel at B = 64 and 4096.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py tests/test_gpu_checkpoint.py -m gpu -q > gpurun_out/pytest_train_$TAG.log 2>&1; echo "train tests rc=$?"; tail -6 gpurun_out/pytest_train_$TAG.log
timeout -s KILL 400 python bench.py --steps 10 --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$TAG.err
for B in 64 4096; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name regex:lbs_tc_kernel --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/lbs_tc_${TAG}_b$B python tools/bench_lbs.py --batches $B --iters 3 --modes tc > gpurun_out/ncu_lbs_${TAG}_b$B.log 2>&1; echo "ncu B=$B rc=$?"
  ncu -i gpurun_out/lbs_tc_${TAG}_b$B.ncu-rep --page raw --csv > gpurun_out/lbs_tc_${TAG}_b${B}_raw.csv 2>/dev/null
  ncu -i gpurun_out/lbs_tc_${TAG}_b$B.ncu-rep --page source --csv > gpurun_out/lbs_tc_${TAG}_b${B}_source.csv 2>/dev/null
done
python tools/show_bench.py gpurun_out/bench_$TAG.json
python - <<PY
import json
d = json.load(open('gpurun_out/bench_$TAG.json'))
t = d.get('train') or {}
print('train eager %s graphed %s note %s launches %s' % (t.get('ms_per_step_eager'), t.get('ms_per_step_graphed'), t.get('graph_note'), t.get('library_launches_per_step')))
PY
