#!/usr/bin/env bash
# ncu --set full capture of the fast kernels (fwd + bwd) of one short bench run.  usage: gpu_ncu.sh TAG [extra bench args]
set -u
mkdir -p gpurun_out
TAG=${1:-p}; shift || true
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $*"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sml_fast|sml_ws|filtergrad" -s 9 -c 3 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu exit $?"; tail -n 2 gpurun_out/ncu_full_$TAG.log
