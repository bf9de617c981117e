#!/bin/bash
# B = 4096 parity test of the tensor-core LBS + ncu --set full of lbs_tc3_kernel at HEAD
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 200 python -m pytest tests/test_gpu_smpl.py -m gpu -q -k "sweep_size or large_batch" > gpurun_out/pytest_lbs4096_$TAG.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_lbs4096_$TAG.log
timeout -s KILL 200 ncu --set full --clock-control none --import-source on --kernel-name regex:lbs_tc3_kernel --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/lbs_tc3_$TAG python tools/bench_lbs.py --batches 4096 --iters 3 --modes tc > gpurun_out/ncu_lbs_tc3_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/lbs_tc3_$TAG.ncu-rep --page raw --csv > gpurun_out/lbs_tc3_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/lbs_tc3_$TAG.ncu-rep --page source --csv > gpurun_out/lbs_tc3_${TAG}_source.csv 2>/dev/null
rm -f gpurun_out/lbs_tc3_$TAG.ncu-rep
python tools/ncu_summary.py gpurun_out/lbs_tc3_${TAG}_raw.csv
