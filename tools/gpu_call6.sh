#!/bin/bash
# ncu --set full of the LBS kernels at B = 4096 (+ chain kernel), then the training-step tests with per-tensor diagnostics.
TAG=${1:-r02}
mkdir -p gpurun_out
for K in lbs_tc3_kernel smpl_chain_kernel joints_kernel; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on --kernel-name regex:$K --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/${K}_${TAG} python tools/bench_lbs.py --batches 4096 --iters 3 --modes tc > gpurun_out/ncu_${K}_${TAG}.log 2>&1; echo "ncu $K rc=$?"
  ncu -i gpurun_out/${K}_${TAG}.ncu-rep --page raw --csv > gpurun_out/${K}_${TAG}_raw.csv 2>/dev/null
  ncu -i gpurun_out/${K}_${TAG}.ncu-rep --page source --csv > gpurun_out/${K}_${TAG}_source.csv 2>/dev/null
  rm -f gpurun_out/${K}_${TAG}.ncu-rep
done
timeout -s KILL 600 python -m pytest tests/test_gpu_train.py tests/test_gpu_train_step.py -m gpu -q -rA > gpurun_out/pytest_train_$TAG.log 2>&1; echo "train tests rc=$?"; grep -n "beyond 0.1 lr\|passed\|failed" gpurun_out/pytest_train_$TAG.log | head -40
