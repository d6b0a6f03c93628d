#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
SML_DEBUG=1 timeout 300 python tools/ws_debug.py 4,2048,768 6,4096,768 5,1024,768 16,8192,768 > gpurun_out/ws_debug4.log 2>&1; echo "ws_debug exit $?"; grep -v "tid=" gpurun_out/ws_debug4.log | cut -c1-200 | tail -n 40
SML_FAST_WS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_ws1.log 2>&1; echo "pytest ws=1 exit $?"; tail -n 4 gpurun_out/pytest_ws1.log
for WS in 1 0; do for dt in f32 bf16; do
SML_FAST_WS=$WS timeout 300 python bench.py --steps 50 --warmup 5 --dtype $dt --no-cpu-baseline --no-e2e > gpurun_out/bench_${dt}_ws$WS.json 2> gpurun_out/bench_${dt}_ws$WS.err; echo "bench ws=$WS $dt exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_${dt}_ws$WS.json"))
    print("ws=$WS $dt", "tok/s %.1fM"%(d["value"]/1e6), "ms/step %.4f"%d["ms_per_step"], "fwd %.4f ms (%.3f)"%(d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"]), "bwd %.4f ms (%.3f)"%(d["roofline"]["launch_ms"], d["roofline"]["frac"]))
except Exception as e: print("no result", e)
PY
done; done
