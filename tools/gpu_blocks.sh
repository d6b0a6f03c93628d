#!/usr/bin/env bash
# GPU check of the block rows (f-1 / f-2 / f-4): parity tests, then block-level timing
set -u
mkdir -p gpurun_out
TAG=${1:-b}
SML_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_blocks.py -q -m gpu -p no:cacheprovider > gpurun_out/pytest_blocks_$TAG.log 2>&1; echo "pytest blocks exit $?"; tail -n 40 gpurun_out/pytest_blocks_$TAG.log
timeout 300 python tools/bench_blocks.py --steps 30 --warmup 5 > gpurun_out/bench_blocks_$TAG.json 2> gpurun_out/bench_blocks_$TAG.err; echo "bench blocks exit $?"; cat gpurun_out/bench_blocks_$TAG.json; tail -n 5 gpurun_out/bench_blocks_$TAG.err
