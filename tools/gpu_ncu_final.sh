#!/usr/bin/env bash
# ncu evidence for the last build: launch list of the default bench command + one --set full capture of fwd, bwd, reduction
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bf16 --no-blocks"
$CMD > gpurun_out/r02f_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02f_ncu_launches.csv $CMD > gpurun_out/r02f_ncu_launches.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:"sml_fast|filtergrad" -s 9 -c 3 -o gpurun_out/r02f_prof_f32 -f $CMD > gpurun_out/r02f_ncu_full.log 2>&1; echo "ncu full exit $?"
