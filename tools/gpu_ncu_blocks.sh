#!/usr/bin/env bash
# ncu captures of the block rows: launch list (durations) + one --set full capture of the extended fused kernel and the LayerNorm kernels
set -u
mkdir -p gpurun_out
TAG=${1:-blk}
CMD="python tools/bench_blocks.py --steps 2 --warmup 2"
$CMD > gpurun_out/plain_blocks_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_blocks_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_blocks_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1; echo "launch list exit $?"
ncu --set full --clock-control none --import-source on -k regex:"sml_fast|ln_stats|ln_backward" -s 8 -c 5 -o gpurun_out/prof_blocks_$TAG -f $CMD > gpurun_out/ncu_full_blocks_$TAG.log 2>&1; echo "ncu full exit $?"; tail -n 2 gpurun_out/ncu_full_blocks_$TAG.log
