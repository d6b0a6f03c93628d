#!/usr/bin/env bash
# A/B of two builds of the library on the same box.  usage: gpu_ab_lib.sh path/to/old.so [reps]
set -u
OLD=$1; REPS=${2:-3}
for i in $(seq $REPS); do for L in "$OLD" ""; do
SML_LIB_PATH=$L timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-e2e --no-bf16 --no-blocks 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('${L:-current}'.split('/')[-1], 'ms/step %.4f fwd %.4f (%.3f) bwd %.4f (%.3f)'%(d['ms_per_step'], d['roofline_fwd']['launch_ms'], d['roofline_fwd']['frac'], d['roofline']['launch_ms'], d['roofline']['frac']))"
done; done
