#!/bin/bash
# in-situ timing experiments for conv_tc_kernel (results are garbage in modes 1-3; only the launch list matters)
# needs the experiment build:  make -C straps-3dhumanshapepose_b200/csrc experiments
export STRAPS_B200_LIB=$PWD/straps-3dhumanshapepose_b200/straps_b200/libstraps_b200_exp.so
for mode in 0 1 2 3; do
  STRAPS_TC_DEBUG=$mode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/conv_exp_$mode.csv \
    python bench.py --steps 1 --warmup 3 --cpu-reps 1 --cpu-sample 2 > /dev/null 2>&1
done
