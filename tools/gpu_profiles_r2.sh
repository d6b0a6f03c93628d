#!/usr/bin/env bash
# round-2 evidence run (1 GPU): full GPU test suite, bench lines, ncu launch list + full captures of the dominant kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -x -p no:cacheprovider > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -n 3 gpurun_out/r02_pytest_gpu.log
python bench.py --steps 200 --warmup 10 > gpurun_out/r02_bench_f32.json 2> gpurun_out/r02_bench_f32.err; echo "bench f32 exit $?"
SML_TC=1 python bench.py --steps 200 --warmup 10 --dtype bf16 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_bf16_tc.json 2> gpurun_out/r02_bench_bf16_tc.err; echo "bench bf16 tc exit $?"
python bench.py --steps 200 --warmup 10 --dtype bf16 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench_bf16.json 2> gpurun_out/r02_bench_bf16.err; echo "bench bf16 exit $?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "reference arm exit $?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-bf16"
$CMD > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_ncu_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:"sml_fast|filtergrad" -s 9 -c 3 -o gpurun_out/r02_prof_f32 -f $CMD > gpurun_out/r02_ncu_full_f32.log 2>&1
echo "ncu full f32 exit $?"
export SML_TC=1
CMD2="python bench.py --steps 2 --warmup 3 --dtype bf16 --no-cpu-baseline --no-e2e"
$CMD2 > gpurun_out/r02_plain_tc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"sml_tc" -s 4 -c 2 -o gpurun_out/r02_prof_tc -f $CMD2 > gpurun_out/r02_ncu_full_tc.log 2>&1
echo "ncu full tc exit $?"
