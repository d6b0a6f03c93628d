#!/usr/bin/env bash
# What the driver runs at round end, in the same order: pytest -m gpu (whole suite, one process), smoke(), reference arm, bench default line.
set -u
mkdir -p gpurun_out
TAG=${1:-full}
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 6 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py --steps ${STEPS:-100} --warmup 10 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench exit $?"; tail -n 3 gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
print("value %.1fM tok/s  ms/step %.4f  step frac %.3f  fwd %.4f (%.3f)  bwd %.4f (%.3f)" % (d["value"]/1e6, d["ms_per_step"], d["roofline_step"]["frac"], d["roofline_fwd"]["launch_ms"], d["roofline_fwd"]["frac"], d["roofline"]["launch_ms"], d["roofline"]["frac"]))
print("bf16", {k: (round(v,4) if isinstance(v,float) else v) for k,v in (d.get("bf16") or {}).items() if k in ("ms_per_step","unavailable")})
print("blocks", json.dumps(d.get("blocks"))[:900])
print("e2e", d.get("e2e")); print("clocks", d.get("clocks"))
PY
