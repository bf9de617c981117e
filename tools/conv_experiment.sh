#!/bin/bash
# In-situ timing experiments for the tensor-core convolution kernels (results are garbage in modes 1-6; only the launch list matters).
# Needs the experiment build:  make -C straps-3dhumanshapepose_b200/csrc experiments
# STRAPS_TC_DEBUG is a bit mask: 1 = no TMA traffic, 2 = no MMAs, 4 = no epilogue global I/O; 3 / 5 / 6 leave ONE part running alone
# (3 = epilogue alone, 5 = MMAs alone, 6 = TMA alone).  VARIANT (optional, "NAME=VALUE ...") adds kernel switches, e.g.
#   tools/conv_experiment.sh "STRAPS_TC_HALO=1"          -> gpurun_out/conv_exp_<mode>_STRAPS_TC_HALO=1.csv
#   tools/conv_experiment.sh "STRAPS_TC_CONV1=s2d"
export STRAPS_B200_LIB=$PWD/straps-3dhumanshapepose_b200/straps_b200/libstraps_b200_exp.so
VARIANT="$1"
TAG=${VARIANT// /_}
for mode in ${MODES:-0 1 2 4 3 5 6}; do
  env $VARIANT STRAPS_TC_DEBUG=$mode timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/conv_exp_${mode}${TAG:+_$TAG}.csv python bench.py --steps 1 --warmup 3 --cpu-reps 1 --cpu-sample 2 > /dev/null 2>&1
done
