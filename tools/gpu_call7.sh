#!/bin/bash
# ncu --set full of the 20 convolution launches of one inference step (roofline.traffic of bench.py) + launch list of a step.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 500 ncu --set full --clock-control none --kernel-name regex:"conv_tc_kernel|conv1_s2d_kernel" --launch-skip 60 --launch-count 20 \
  -f -o gpurun_out/conv_full_$TAG python bench.py --steps 1 --warmup 3 --train-steps 0 --no-graph --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 > gpurun_out/ncu_conv_$TAG.log 2>&1; echo "ncu conv rc=$?"
ncu -i gpurun_out/conv_full_$TAG.ncu-rep --page raw --csv > gpurun_out/conv_full_${TAG}_raw.csv 2>/dev/null
rm -f gpurun_out/conv_full_$TAG.ncu-rep
python tools/ncu_summary.py gpurun_out/conv_full_${TAG}_raw.csv --traffic-json gpurun_out/conv_traffic_$TAG.json \
  --source "ncu --set full --clock-control none, the 20 convolution launches (conv1_s2d_kernel + 19 conv_tc_kernel) of one inference step, bench.py --steps 1 --warmup 3, B=64, C=17 ($TAG)" | tee gpurun_out/conv_full_${TAG}_summary.txt
