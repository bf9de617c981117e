#!/bin/bash
# One gpurun call: GPU test suite, the bench line, and an ncu launch list (inference + training step) of the same command.
#   gpurun --timeout 1500 -- 'bash tools/gpu_call.sh [tag]'
TAG=${1:-r02}
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$TAG.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --train-steps 2 --no-train-graph --no-graph --no-lbs-sweep --cpu-reps 1 --cpu-sample 2 > gpurun_out/ncu_bench_$TAG.log 2>&1; echo "ncu rc=$?"
python tools/show_bench.py gpurun_out/bench_$TAG.json
