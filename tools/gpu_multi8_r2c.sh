#!/usr/bin/env bash
# N-GPU validation after pass splitting + block rows: what the driver's scaling run does (cfg-2 weak scaling, default collective),
# cfg-3 strong scaling, the long-context sweep with and without pass splitting
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --no-blocks > gpurun_out/r2c_cfg2_N$N.json 2> gpurun_out/r2c_cfg2_N$N.err; echo "cfg2 N=$N exit $?"
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --config cfg3 --no-e2e > gpurun_out/r2c_cfg3_N$N.json 2> gpurun_out/r2c_cfg3_N$N.err; echo "cfg3 N=$N exit $?"
timeout 900 $TR --master-port 29515 bench.py --gpus $N --config cfg5 --steps 10 > gpurun_out/r2c_cfg5_N$N.json 2> gpurun_out/r2c_cfg5_N$N.err; echo "cfg5 N=$N exit $?"
SML_SPLIT=1 timeout 900 $TR --master-port 29516 bench.py --gpus $N --config cfg5 --steps 10 > gpurun_out/r2c_cfg5_nosplit_N$N.json 2> gpurun_out/r2c_cfg5_nosplit_N$N.err; echo "cfg5 (SML_SPLIT=1) N=$N exit $?"
python - <<PY
import json
for f in ("r2c_cfg2_N$N", "r2c_cfg3_N$N", "r2c_cfg5_N$N", "r2c_cfg5_nosplit_N$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], (d.get("impl_detail") or {}).get("collective"), d.get("e2e") and "e2e %.2fM" % (d["e2e"]["value"] / 1e6), d.get("bf16") and "bf16 %.4f ms" % d["bf16"].get("ms_per_step", -1))
        for r in d.get("sweep", []):
            print("   T=%d B/gpu=%d %.4f ms frac %.3f" % (r["seq_len"], r["batch_per_gpu"], r["ms_per_step"], r["roofline_step_frac"]))
    except Exception as e:
        print(f, "no result", e)
PY
