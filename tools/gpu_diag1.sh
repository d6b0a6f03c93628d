#!/usr/bin/env bash
# diagnostics: (1) warp-specialised kernel on multi-tile shapes with the mbarrier-timeout record, (2) tile quantisation A/B
set -u
mkdir -p gpurun_out
SML_DEBUG=1 timeout 300 python tools/ws_debug.py > gpurun_out/ws_debug.log 2>&1; echo "ws_debug exit $?"; tail -n 30 gpurun_out/ws_debug.log
for BATCH in 16 37 32; do
timeout 300 python bench.py --steps 20 --warmup 5 --batch $BATCH --no-cpu-baseline --no-e2e > gpurun_out/bench_f32_B$BATCH.json 2> gpurun_out/bench_f32_B$BATCH.err; echo "bench B=$BATCH exit $?"
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_f32_B$BATCH.json"))
    print("B=$BATCH", "tok/s %.1fM"%(d["value"]/1e6), "ms/step %.4f"%d["ms_per_step"], "fwd %.4f ms (%.3f)"%(d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"]), "bwd %.4f ms (%.3f)"%(d["roofline"]["launch_ms"], d["roofline"]["frac"]))
except Exception as e: print("no result", e)
PY
done
