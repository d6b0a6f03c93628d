#!/usr/bin/env bash
# one vs two TMA landing tiles over a few shapes / dtypes (SML_FAST_XB), with SML_FAST_CTAS=2 for the fp32 M=1024 kernel
set -u
mkdir -p gpurun_out
run() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e "$@" > gpurun_out/ab_xb.json 2> gpurun_out/ab_xb.err || tail -n 3 gpurun_out/ab_xb.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab_xb.json"))
    print("$label", "tok/s %.1fM"%(d["value"]/1e6), "ms/step %.4f"%d["ms_per_step"], "fwd %.4f ms (%.3f)"%(d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"]), "bwd %.4f ms (%.3f)"%(d["roofline"]["launch_ms"], d["roofline"]["frac"]))
except Exception as e: print("$label no result", e)
PY
}
for XB in 1 2; do
run "cfg2 bf16 xb=$XB" SML_FAST_XB=$XB -- --dtype bf16
run "cfg2 bf16 xb=$XB (repeat)" SML_FAST_XB=$XB -- --dtype bf16
run "cfg3 f32 B=64 xb=$XB" SML_FAST_XB=$XB -- --batch 64 --seq 4096 --embed 1024
run "cfg3 bf16 B=64 xb=$XB" SML_FAST_XB=$XB -- --batch 64 --seq 4096 --embed 1024 --dtype bf16
run "cfg1 f32 xb=$XB" SML_FAST_XB=$XB -- --batch 8 --seq 512 --embed 256
run "D256 T8192 B32 f32 xb=$XB" SML_FAST_XB=$XB -- --batch 32 --seq 8192 --embed 256
done
run "cfg2 f32 ctas=3 xb=1" SML_FAST_XB=1 --
run "cfg2 f32 ctas=2 xb=2" SML_FAST_XB=2 SML_FAST_CTAS=2 --
