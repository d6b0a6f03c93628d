#!/usr/bin/env bash
# run before every gpurun: build + CPU test suite (catches missing exports / syntax errors for free)
set -e
cd "$(dirname "$0")/.."
make -j8 -C tensor-cuda-fft-_b200/csrc 2>&1 | grep -E "error|Error" -A5 && exit 1
python -m pytest tests -x -q -m "not gpu" > /tmp/preflight_pytest.log 2>&1 || { tail -30 /tmp/preflight_pytest.log; exit 1; }
tail -1 /tmp/preflight_pytest.log
