#!/usr/bin/env bash
# configs[2] (cfg3, strong scaling) and configs[4] (cfg5, long-context sweep) at N GPUs; at N = 1 the sweep carries the CPU column
set -u
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --config cfg5 --steps 10 --warmup 3 > gpurun_out/r2d_cfg5_N1.json 2> gpurun_out/r2d_cfg5_N1.err; echo "cfg5 N=1 (with CPU column) exit $?"
  timeout 600 python bench.py --config cfg3 --steps 100 --warmup 10 --no-e2e > gpurun_out/r2d_cfg3_N1.json 2> gpurun_out/r2d_cfg3_N1.err; echo "cfg3 N=1 exit $?"
else
  TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
  timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 --config cfg3 --no-e2e > gpurun_out/r2d_cfg3_N$N.json 2> gpurun_out/r2d_cfg3_N$N.err; echo "cfg3 N=$N exit $?"
  timeout 900 $TR --master-port 29515 bench.py --gpus $N --config cfg5 --steps 10 > gpurun_out/r2d_cfg5_N$N.json 2> gpurun_out/r2d_cfg5_N$N.err; echo "cfg5 N=$N exit $?"
fi
python - <<PY
import json
for f in ("r2d_cfg3_N$N", "r2d_cfg5_N$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], (d.get("impl_detail") or {}).get("collective"))
        for r in d.get("sweep", []):
            c = r.get("cpu_reference")
            print("   T=%d B/gpu=%d %.4f ms frac %.3f %s" % (r["seq_len"], r["batch_per_gpu"], r["ms_per_step"], r["roofline_step_frac"], ("cpu %.1fk tok/s" % (c["value"] / 1e3)) if c else ""))
    except Exception as e:
        print(f, "no result", e)
PY
