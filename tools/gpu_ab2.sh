#!/usr/bin/env bash
# A/B over an env knob, parity tests under every value.  usage: gpu_ab2.sh VAR "v1 v2 ..." [dtypes]
set -u
mkdir -p gpurun_out
VAR=${1:-SML_FAST_P}; VALS=${2:-"4 6"}; DTS=${3:-"f32 bf16"}
for V in $VALS; do
env $VAR=$V timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_${VAR}_$V.log 2>&1; echo "$VAR=$V pytest exit $?"; tail -n 3 gpurun_out/pytest_${VAR}_$V.log
for dt in $DTS; do
env $VAR=$V timeout 600 python bench.py --steps 50 --warmup 5 --dtype $dt --no-cpu-baseline --no-e2e > gpurun_out/bench_${dt}_${VAR}_$V.json 2> gpurun_out/bench_${dt}_${VAR}_$V.err; echo "bench $VAR=$V $dt exit $?"; tail -n 3 gpurun_out/bench_${dt}_${VAR}_$V.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${dt}_${VAR}_$V.json"))
    print("$VAR=$V $dt", "tok/s %.1fM"%(d["value"]/1e6), "ms/step %.4f"%d["ms_per_step"], "fwd %.4f ms (%.3f)"%(d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"]), "bwd %.4f ms (%.3f)"%(d["roofline"]["launch_ms"], d["roofline"]["frac"]))
except Exception as e: print("no result", e)
PY
done; done
