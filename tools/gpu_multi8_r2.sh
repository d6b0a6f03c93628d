#!/usr/bin/env bash
# 8-GPU (or N-GPU) validation of round 2: fused-collective parity, cfg-2 weak scaling fused vs NCCL (+ e2e), cfg-3 strong scaling,
# long-context sweep, host-link ceiling with every rank copying at once
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29520 tools/fused_allreduce_check.py 2>&1 | grep "shape" | tail -9
for mode in fused nccl; do
  SML_ALLREDUCE=$mode timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 $([ $mode != fused ] && echo "--no-e2e --no-bf16") > gpurun_out/r2_cfg2_${mode}_N$N.json 2> gpurun_out/r2_cfg2_${mode}_N$N.err; echo "cfg2 $mode N=$N exit $?"
  SML_ALLREDUCE=$mode timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --config cfg3 --no-e2e > gpurun_out/r2_cfg3_${mode}_N$N.json 2> gpurun_out/r2_cfg3_${mode}_N$N.err; echo "cfg3 $mode N=$N exit $?"
done
timeout 900 $TR --master-port 29515 bench.py --gpus $N --config cfg5 --steps 10 > gpurun_out/r2_cfg5_N$N.json 2> gpurun_out/r2_cfg5_N$N.err; echo "cfg5 N=$N exit $?"
timeout 300 $TR --master-port 29513 tools/pcie_bw.py > gpurun_out/r2_pcie_N$N.txt 2>&1; grep rank gpurun_out/r2_pcie_N$N.txt | cut -c1-250
python - <<PY
import json
for f in ("r2_cfg2_fused_N$N", "r2_cfg2_nccl_N$N", "r2_cfg3_fused_N$N", "r2_cfg3_nccl_N$N", "r2_cfg5_N$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], (d.get("impl_detail") or {}).get("collective"), d.get("e2e") and "e2e %.2fM" % (d["e2e"]["value"] / 1e6), d.get("bf16") and "bf16 %.4f ms" % d["bf16"].get("ms_per_step", -1))
        for r in d.get("sweep", []):
            print("   T=%d B/gpu=%d %.4f ms frac %.3f" % (r["seq_len"], r["batch_per_gpu"], r["ms_per_step"], r["roofline_step_frac"]))
    except Exception as e:
        print(f, "no result", e)
PY
