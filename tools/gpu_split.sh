#!/usr/bin/env bash
# pass splitting: parity (whole GPU suite exercises it on the small shapes), cfg-2 must not move, long-context sweep with and without it
set -u
mkdir -p gpurun_out
TAG=${1:-s}
timeout 1500 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_split_$TAG.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 3 gpurun_out/pytest_split_$TAG.log
for sp in 1 0; do
  SML_SPLIT=$sp timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-bf16 --no-blocks > gpurun_out/split_cfg2_$sp.json 2> gpurun_out/split_cfg2_$sp.err; echo "cfg2 SML_SPLIT=$sp exit $?"
  SML_SPLIT=$sp timeout 600 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/split_cfg5_$sp.json 2> gpurun_out/split_cfg5_$sp.err; echo "cfg5 SML_SPLIT=$sp exit $?"
done
# one batch element per rank (what N = 8 sees at T = 128K / 64K), on one GPU
for sp in 1 0; do
  for T in 131072 65536 32768; do
    SML_SPLIT=$sp timeout 300 python bench.py --batch 1 --seq $T --embed 1024 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-bf16 --no-blocks > gpurun_out/split_b1_${T}_$sp.json 2> gpurun_out/split_b1_${T}_$sp.err; echo "B=1 T=$T SML_SPLIT=$sp exit $?"
  done
done
python - <<'PY'
import json
for sp in (1, 0):
    d = json.load(open(f"gpurun_out/split_cfg2_{sp}.json"))
    print("cfg2 split", sp, "ms/step %.4f fwd %.4f bwd %.4f" % (d["ms_per_step"], d["roofline_fwd"]["launch_ms"], d["roofline"]["launch_ms"]))
    d = json.load(open(f"gpurun_out/split_cfg5_{sp}.json"))
    print("cfg5 split", sp, " ".join("%dK:%.2f(%.2f)" % (r["seq_len"] // 1024, r["ms_per_step"], r["roofline_step_frac"]) for r in d["sweep"]))
    for T in (131072, 65536, 32768):
        d = json.load(open(f"gpurun_out/split_b1_{T}_{sp}.json"))
        print("B=1 T=%d split %d ms/step %.4f step frac %.3f" % (T, sp, d["ms_per_step"], d["roofline_step"]["frac"]))
PY
