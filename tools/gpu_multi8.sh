#!/usr/bin/env bash
# N-GPU validation: cfg-2 weak scaling, cfg-3 (global batch 64 sharded over N), LM training step under DDP
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_f32_N$N.json 2> gpurun_out/bench_f32_N$N.err; echo "bench cfg2 N=$N exit $?"; tail -n 3 gpurun_out/bench_f32_N$N.err | cut -c1-300
timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 100 --warmup 10 --batch $((64 / N)) --seq 4096 --embed 1024 --no-e2e > gpurun_out/bench_cfg3_N$N.json 2> gpurun_out/bench_cfg3_N$N.err; echo "bench cfg3 N=$N exit $?"; tail -n 3 gpurun_out/bench_cfg3_N$N.err | cut -c1-300
timeout 600 $TR --master-port 29513 tools/lm_train_step.py > gpurun_out/lm_N$N.json 2> gpurun_out/lm_N$N.err; echo "lm N=$N exit $?"; tail -n 3 gpurun_out/lm_N$N.err | cut -c1-300
python - <<PY
import json
for f in ("bench_f32_N$N", "bench_cfg3_N$N", "lm_N$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], d.get("e2e") and "e2e %.2fM" % (d["e2e"]["value"] / 1e6))
    except Exception as e:
        print(f, "no result", e)
PY
