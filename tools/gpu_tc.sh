#!/usr/bin/env bash
# tensor-core kernel: intermediate dumps of one work item vs the numpy model, parity + timing per shape, phase timeline of cfg-2
set -u
mkdir -p gpurun_out
{
for a in "512 16" "8192 384"; do echo "--- tc_dump_check $a"; timeout 300 python tools/tc_dump_check.py $a; echo "exit $?"; done
SML_NO_DEBUG=1 timeout 1500 python tools/tc_check.py ${1:-}; echo "tc_check exit $?"
echo "--- timeline (SML_DEBUG=1: timing build of the forward kernel)"
SML_TC=1 SML_DEBUG=1 python tools/tc_check.py 16 8192 768 384 2>&1 | grep "timeline"
} > gpurun_out/tc_check.log 2>&1
grep -v "bad \|worst" gpurun_out/tc_check.log | cut -c1-330
