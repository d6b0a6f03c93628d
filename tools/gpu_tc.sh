#!/usr/bin/env bash
# tensor-core kernel bring-up: intermediate dumps of one work item vs the numpy model, then parity + timing per shape
set -u
mkdir -p gpurun_out
{
for a in "512 16" "8192 384"; do echo "--- tc_dump_check $a"; timeout 300 python tools/tc_dump_check.py $a; echo "exit $?"; done
timeout 1500 python tools/tc_check.py ${1:-}; echo "tc_check exit $?"
} > gpurun_out/tc_check.log 2>&1
grep -v "bad \|worst" gpurun_out/tc_check.log | cut -c1-330
