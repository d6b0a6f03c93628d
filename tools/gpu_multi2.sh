#!/usr/bin/env bash
# N-GPU validation of round 2: symmetric-memory gradient bucket vs NCCL at cfg-2 (weak) and cfg-3 (strong), host-link ceiling with all ranks copying
set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for mode in fused symm nccl; do
  SML_ALLREDUCE=$mode timeout 600 $TR --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --no-bf16 $([ $mode != fused ] && echo --no-e2e) > gpurun_out/r2_cfg2_${mode}_N$N.json 2> gpurun_out/r2_cfg2_${mode}_N$N.err; echo "cfg2 $mode N=$N exit $?"; tail -n 2 gpurun_out/r2_cfg2_${mode}_N$N.err | cut -c1-300
  SML_ALLREDUCE=$mode timeout 600 $TR --master-port 29512 bench.py --gpus $N --steps 200 --warmup 10 --config cfg3 --no-e2e > gpurun_out/r2_cfg3_${mode}_N$N.json 2> gpurun_out/r2_cfg3_${mode}_N$N.err; echo "cfg3 $mode N=$N exit $?"; tail -n 2 gpurun_out/r2_cfg3_${mode}_N$N.err | cut -c1-300
done
timeout 300 $TR --master-port 29513 tools/pcie_bw.py > gpurun_out/r2_pcie_N$N.txt 2>&1; cat gpurun_out/r2_pcie_N$N.txt | grep rank
timeout 300 $TR --master-port 29514 tools/pcie_bw.py --no-bind > gpurun_out/r2_pcie_nobind_N$N.txt 2>&1; cat gpurun_out/r2_pcie_nobind_N$N.txt | grep rank
python - <<PY
import json
for f in ("r2_cfg2_fused_N$N", "r2_cfg2_symm_N$N", "r2_cfg2_nccl_N$N", "r2_cfg3_fused_N$N", "r2_cfg3_symm_N$N", "r2_cfg3_nccl_N$N"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value %.1fM" % (d["value"] / 1e6), "ms/step %.4f" % d["ms_per_step"], d["impl_detail"]["collective"], d.get("e2e") and "e2e %.2fM %s" % (d["e2e"]["value"] / 1e6, d["e2e"].get("host_affinity")))
    except Exception as e:
        print(f, "no result", e)
PY
