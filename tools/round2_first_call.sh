#!/bin/bash
# First gpurun call of round 2: every kernel that was written after round 1's GPU budget was spent, against the shipped path.
#   gpurun --timeout 900 -- 'bash tools/round2_first_call.sh'        (about 6-8 minutes of box time: 4 check runs of ~10 s, the parked
#   tests ~2 min, 8 ncu launch lists of ~30 s)
# 1. tools/halo_check.py: errors (features + every layer activation) and encoder timings of each variant, one process, JSON flushed after
#    every stage (a hang leaves the name of the variant under "reached"); each run under its own timeout so that a hung kernel costs
#    60 s, not the call.  Order: the conv1 pair-layout kernels (biggest predicted gain), the merged CTA-pair kernel, the combinations.
# 2. the same with dense 96-byte pair lines (STRAPS_TC_S2D_PITCH=48: does TMA accept a 96-byte inner box under SWIZZLE_128B?).
# 3. the parked GPU tests.
# 4. component-alone timings (experiment build) of the shipped path and of the halo kernel: DESIGN.md 4.2 item 5 (a) vs (b).
mkdir -p gpurun_out
run() { name=$1; shift; timeout -s KILL 120 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; }
run check_s2d    python tools/halo_check.py --out gpurun_out/check_s2d.json --only conv1_s2d,conv1_s2d2,conv1_s2dp
run check_pair   python tools/halo_check.py --out gpurun_out/check_pair.json --only pair,pair_m128,pair_m
run check_combo  python tools/halo_check.py --out gpurun_out/check_combo.json --only pair_m+s2dp,tma2,pdl,pdl+s2dp,halo,halo2
STRAPS_TC_S2D_PITCH=48 run check_s2d_p48 python tools/halo_check.py --out gpurun_out/check_s2d_p48.json --only conv1_s2d,conv1_s2d2,conv1_s2dp
STRAPS_TEST_UNVERIFIED=1 timeout -s KILL 600 python -m pytest tests/test_gpu_zz_conv_kernels.py tests/test_gpu_conv_variants.py -q -m gpu > gpurun_out/pytest_unverified.log 2>&1
echo "pytest rc=$?"
MODES="0 3 5 6" timeout -s KILL 400 bash tools/conv_experiment.sh
MODES="0 3 5 6" timeout -s KILL 400 bash tools/conv_experiment.sh "STRAPS_TC_HALO=1"
tail -c 600 gpurun_out/check_s2d.json gpurun_out/check_pair.json gpurun_out/check_combo.json
