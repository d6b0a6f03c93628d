#!/usr/bin/env bash
# runs every umma_probe test/variant in its own process (a bad descriptor can fault the context)
set -u
mkdir -p gpurun_out
P=tools/microbench/umma_probe
{
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
for tv in "1 0" "1 1" "2 0" "2 1" "3 0" "3 1" "3 2" "3 3" "3 4" "3 5" "4 0" "4 1" "5 0" "6 0"; do
  echo "--- umma_probe $tv"
  timeout 60 $P $tv; echo "exit $?"
done
} > gpurun_out/umma_probe.log 2>&1
cat gpurun_out/umma_probe.log
