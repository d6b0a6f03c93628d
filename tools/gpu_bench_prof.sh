#!/usr/bin/env bash
# bench (f32 + bf16) and the two ncu passes of B200_PROFILING.md.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_f32_$TAG.json 2> gpurun_out/bench_f32_$TAG.err; echo "bench f32 exit $?"; cat gpurun_out/bench_f32_$TAG.json; tail -n 5 gpurun_out/bench_f32_$TAG.err
timeout 600 python bench.py --steps 200 --warmup 10 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16_$TAG.json 2> gpurun_out/bench_bf16_$TAG.err; echo "bench bf16 exit $?"; cat gpurun_out/bench_bf16_$TAG.json; tail -n 5 gpurun_out/bench_bf16_$TAG.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>&1; cat gpurun_out/bench_ref_$TAG.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launches_$TAG.log 2>&1
echo "ncu launches exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sml_fast|sml_ws|filtergrad" -s 9 -c 3 -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "ncu full exit $?"; tail -n 3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
