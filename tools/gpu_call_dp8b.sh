#!/bin/bash
# 8-GPU A/B of the training arm: all-reduce between the graphs (STRAPS_DP_OVERLAP=0) vs layer4 + IEF part started inside the backward pass
N=${1:-8}
TAG=${2:-r02}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for OV in 0 1 0 1; do
  STRAPS_DP_OVERLAP=$OV timeout -s KILL 200 $RUN bench.py --gpus $N --train-only --train-steps 20 2> gpurun_out/train_${N}gpu_ov${OV}_$TAG.err | grep '^{' >> gpurun_out/train_${N}gpu_$TAG.jsonl; echo "overlap=$OV rc=$?"
done
python - <<PY
import json
for l in open('gpurun_out/train_${N}gpu_$TAG.jsonl'):
    t = json.loads(l)['train']
    print('eager %.3f graphed %s  identical %s  %s' % (t['ms_per_step_eager'], t['ms_per_step_graphed'], t['replicas_identical_after_update'], t['config'][-120:]))
PY
