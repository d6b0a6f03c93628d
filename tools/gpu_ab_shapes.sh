#!/usr/bin/env bash
# lockstep vs warp-specialised kernel over a few shapes (fp32): calibrates the planner's choice
set -u
mkdir -p gpurun_out
for SHAPE in "64 4096 1024" "37 8192 768" "16 8192 768" "8 8192 768" "4 32768 1024" "16 2048 768"; do
set -- $SHAPE
for WS in 0 1; do
SML_FAST_WS=$WS timeout 300 python bench.py --steps 30 --warmup 5 --batch $1 --seq $2 --embed $3 --no-cpu-baseline --no-e2e > gpurun_out/ab_shape.json 2> gpurun_out/ab_shape.err || tail -n 3 gpurun_out/ab_shape.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_shape.json"))
    print("B,T,D=$SHAPE ws=$WS", "tok/s %.1fM"%(d["value"]/1e6), "ms/step %.4f"%d["ms_per_step"], "fwd %.4f ms (%.3f)"%(d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"]), "bwd %.4f ms (%.3f)"%(d["roofline"]["launch_ms"], d["roofline"]["frac"]))
except Exception as e: print("no result", e)
PY
done; done
