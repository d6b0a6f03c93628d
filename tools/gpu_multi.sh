#!/usr/bin/env bash
# multi-GPU bench (one process per GPU over NCCL) + the reference arm launched the same way.  usage: gpu_multi.sh N
set -u
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/bench_f32_N$N.json 2> gpurun_out/bench_f32_N$N.err; echo "bench N=$N exit $?"; tail -n 5 gpurun_out/bench_f32_N$N.err; cat gpurun_out/bench_f32_N$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_ref_N$N.json 2> gpurun_out/bench_ref_N$N.err; echo "reference arm N=$N exit $?"; cat gpurun_out/bench_ref_N$N.json | cut -c1-300
