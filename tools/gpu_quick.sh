#!/usr/bin/env bash
# quick GPU check: all parity tests in one process + f32/bf16 bench without CPU baseline
set -u
mkdir -p gpurun_out
TAG=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest exit $?"; tail -n 8 gpurun_out/pytest_$TAG.log
for dt in f32 bf16; do
timeout 600 python bench.py --steps 20 --warmup 5 --dtype $dt --no-cpu-baseline --no-e2e > gpurun_out/bench_${dt}_$TAG.json 2> gpurun_out/bench_${dt}_$TAG.err; echo "bench $dt exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_${dt}_$TAG.json"))
print("$dt", "tok/s %.1fM"%(d["value"]/1e6), "ms/step %.4f"%d["ms_per_step"], "fwd %.4f ms (%.3f)"%(d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"]), "bwd %.4f ms (%.3f)"%(d["roofline"]["launch_ms"], d["roofline"]["frac"]), d["clocks"])
PY
done
