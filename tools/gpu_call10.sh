#!/bin/bash
# slab epilogue as the default (fp32 outputs included): conv / training parity tests, then the bench line with train arm, direct vs slab
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_gpu_regressor.py tests/test_gpu_numeric_range.py tests/test_gpu_stem_kernels.py tests/test_gpu_train.py tests/test_gpu_train_step.py -m gpu -q > gpurun_out/pytest_slab_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/pytest_slab_$TAG.log
for MODE in direct slab; do
  if [ $MODE = direct ]; then export STRAPS_TC_EPI=direct; else unset STRAPS_TC_EPI; fi
  timeout -s KILL 300 python bench.py --steps 30 --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 2> /dev/null | grep '^{' > gpurun_out/bench_${MODE}_$TAG.json
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_${MODE}_$TAG.json'))
t = d['train']
print('$MODE: ms/step %.4f (eager %.4f) encoder %.4f frac %.4f | train eager %.3f graphed %s' % (d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['encoder_ms'], d['roofline']['frac'], t['ms_per_step_eager'], t['ms_per_step_graphed']))
PY
done
