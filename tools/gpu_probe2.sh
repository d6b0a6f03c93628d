#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
{
for tv in "5 0" "5 1" "5 2" "5 3" "7 0" "7 1"; do echo "--- umma_probe $tv"; timeout 60 tools/microbench/umma_probe $tv; echo "exit $?"; done
} > gpurun_out/umma_probe2.log 2>&1
cat gpurun_out/umma_probe2.log
