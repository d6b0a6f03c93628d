#!/usr/bin/env bash
# Runs on the GPU box (via gpurun): GPU parity tests in isolated processes, smoke, bench.  Logs -> gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
GROUPS_K=("golden or known_answer" "random_shapes" "bf16" "nonlearnable or accumulates or state_dict or dropout" "c_abi" "wirtinger" "full_size" "long_context")
i=0
for k in "${GROUPS_K[@]}"; do
  i=$((i+1))
  timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "$k" -p no:cacheprovider > gpurun_out/pytest_$i.log 2>&1
  echo "group $i [$k] exit $?" | tee -a gpurun_out/summary.txt
  tail -n 25 gpurun_out/pytest_$i.log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -n 5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; echo "bench f32 exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench_f32.json; tail -n 5 gpurun_out/bench_f32.err
timeout 600 python bench.py --steps 10 --warmup 3 --dtype bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench_bf16.json
cat gpurun_out/summary.txt
