#!/usr/bin/env bash
# N-GPU: fused (multimem.red.v4) vs symm (torch multimem all-reduce in place) vs NCCL at cfg-2 weak scaling
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29520 tools/fused_allreduce_check.py 2>&1 | grep "shape" | tail -9 | cut -c1-200
for mode in fused symm nccl; do
  SML_ALLREDUCE=$mode timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 300 --warmup 10 --no-e2e --no-bf16 > gpurun_out/r2b_cfg2_${mode}_N$N.json 2> gpurun_out/r2b_cfg2_${mode}_N$N.err; echo "cfg2 $mode N=$N exit $?"
done
python - <<PY
import json
for f in ("r2b_cfg2_fused_N$N", "r2b_cfg2_symm_N$N", "r2b_cfg2_nccl_N$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], (d.get("impl_detail") or {}).get("collective"))
    except Exception as e:
        print(f, "no result", e)
PY
