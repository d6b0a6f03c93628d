#!/usr/bin/env bash
# ncu --set full capture of the tensor-core kernels (fwd + bwd) of one short bf16 bench run.  usage: gpu_ncu_tc.sh TAG
set -u
mkdir -p gpurun_out
TAG=${1:-tc}
export SML_TC=1
CMD="python bench.py --steps 2 --warmup 3 --dtype bf16 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sml_tc" -s 4 -c 2 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_full_$TAG.log; cat gpurun_out/plain_$TAG.log | cut -c1-300
