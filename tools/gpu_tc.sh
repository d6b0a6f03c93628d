#!/usr/bin/env bash
# tensor-core kernel bring-up: TMA swizzle probe, intermediate dumps of one work item vs the numpy model, then parity + timing per shape
set -u
mkdir -p gpurun_out
{
for v in 0 2; do echo "--- umma_probe 5 $v"; timeout 60 tools/microbench/umma_probe 5 $v; echo "exit $?"; done
for a in "512 16" "512 256" "2048 48" "8192 384"; do echo "--- tc_dump_check $a"; timeout 300 python tools/tc_dump_check.py $a; echo "exit $?"; done
timeout 1500 python tools/tc_check.py; echo "tc_check exit $?"
} > gpurun_out/tc_check.log 2>&1
cat gpurun_out/tc_check.log
