#!/bin/bash
# A/B of the slab epilogue of conv_tc_kernel (STRAPS_TC_EPI=slab): parity tests under the switch, then the bench line with and without it
TAG=${1:-r02}
mkdir -p gpurun_out
STRAPS_TC_EPI=slab timeout -s KILL 300 python -m pytest tests/test_gpu_regressor.py tests/test_gpu_numeric_range.py tests/test_gpu_stem_kernels.py -m gpu -q > gpurun_out/pytest_slab_$TAG.log 2>&1; echo "slab tests rc=$?"; tail -3 gpurun_out/pytest_slab_$TAG.log
for MODE in off slab off slab; do
  if [ $MODE = slab ]; then export STRAPS_TC_EPI=slab; else unset STRAPS_TC_EPI; fi
  timeout -s KILL 200 python bench.py --steps 30 --train-steps 0 --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 2> /dev/null | grep '^{' > gpurun_out/bench_${MODE}_$TAG.json
  python - <<PY
import json
d = json.load(open('gpurun_out/bench_${MODE}_$TAG.json'))
print('$MODE: ms/step %.4f (eager %.4f) encoder %.4f frac %.4f' % (d['ms_per_step'], d['ms_per_step_eager'], d['roofline']['encoder_ms'], d['roofline']['frac']))
PY
done
