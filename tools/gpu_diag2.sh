#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SML_DEBUG=1 timeout 300 python tools/ws_debug.py 4,2048,768 6,4096,768 16,8192,768 > gpurun_out/ws_debug2.log 2>&1; echo "ws_debug exit $?"; tail -n 60 gpurun_out/ws_debug2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider --durations=8 > gpurun_out/pytest_diag2.log 2>&1; echo "pytest exit $?"; tail -n 16 gpurun_out/pytest_diag2.log
for dt in f32 bf16; do
timeout 300 python bench.py --steps 20 --warmup 5 --dtype $dt --no-cpu-baseline > gpurun_out/bench_${dt}_d2.json 2> gpurun_out/bench_${dt}_d2.err; echo "bench $dt exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${dt}_d2.json"))
    print("$dt", "e2e %.2fM tok/s %.2f ms"%(d["e2e"]["value"]/1e6, d["e2e"]["ms_per_step"]), "tok/s %.1fM"%(d["value"]/1e6), "ms/step %.4f"%d["ms_per_step"], "fwd %.4f ms (%.3f)"%(d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"]), "bwd %.4f ms (%.3f)"%(d["roofline"]["launch_ms"], d["roofline"]["frac"]))
except Exception as e: print("no result", e)
PY
done
