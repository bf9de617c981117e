#!/bin/bash
# N-GPU call (gpurun --gpus N): config 4 under torchrun -- the NCCL correctness check of the sharded training step, then the bench line.
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_call_dp.sh 2 r02'
N=${1:-2}
TAG=${2:-r02}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu_$TAG.txt 2>&1
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout -s KILL 400 $RUN tools/dp_check.py --batch 64 --steps 5 --overlap > gpurun_out/dp_check_${N}gpu_$TAG.json 2> gpurun_out/dp_check_${N}gpu_$TAG.err; echo "dp_check rc=$?"; tail -c 600 gpurun_out/dp_check_${N}gpu_$TAG.json
timeout -s KILL 600 $RUN bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu_$TAG.json 2> gpurun_out/bench_${N}gpu_$TAG.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_${N}gpu_$TAG.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_${N}gpu_$TAG.json'))
t = d.get('train') or {}
print('N=%d value %.0f e2e %.0f kp %.0f' % (d['n_gpus'], d['value'], d['e2e']['value'], (d.get('e2e_from_keypoints') or {}).get('value', 0)))
print('h2d', d['e2e'].get('h2d_copy_alone_gbs_per_rank'), d['e2e'].get('numa'))
print('train', {k: t.get(k) for k in ('ms_per_step', 'ms_per_step_eager', 'ms_per_step_graphed', 'value', 'replicas_identical_after_update', 'graph_note', 'allreduce_elements')})
PY
