#!/bin/bash
# 8-GPU call (charged 8x): topology + the bench line under torchrun (inference replicas, e2e with the H2D ceiling probe, config-4 training arm)
N=${1:-8}
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu_$TAG.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" >> gpurun_out/topo_${N}gpu_$TAG.txt 2>&1
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 5 \
  > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
d = json.loads([l for l in open('gpurun_out/bench_${N}gpu_$TAG.json') if l.startswith('{')][-1])
t = d.get('train') or {}
print('N=%d value %.0f e2e %.0f kp %.0f' % (d['n_gpus'], d['value'], d['e2e']['value'], (d.get('e2e_from_keypoints') or {}).get('value', 0)))
print('h2d', d['e2e'].get('h2d_copy_alone_gbs_per_rank'), d['e2e'].get('numa'))
print('train', {k: t.get(k) for k in ('ms_per_step', 'ms_per_step_eager', 'ms_per_step_graphed', 'value', 'replicas_identical_after_update', 'graph_note')})
PY
