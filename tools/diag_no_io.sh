#!/usr/bin/env bash
# Diagnostic builds of the library without TMA loads / stores (results are garbage; only the timing of the SM side matters).
# usage (CPU box): tools/diag_no_io.sh   -> tensor-cuda-fft-_b200/libsml_diag_{noload,nostore,noio}.so
set -eu
cd "$(dirname "$0")/../tensor-cuda-fft-_b200/csrc"
for v in noload nostore noio; do
  case $v in noload) F="-DSML_DIAG_NO_LOAD";; nostore) F="-DSML_DIAG_NO_STORE";; noio) F="-DSML_DIAG_NO_LOAD -DSML_DIAG_NO_STORE";; esac
  mkdir -p build_$v
  for f in sml_api sml_inst_f32_fwd sml_inst_f32_bwd sml_inst_bf16_fwd sml_inst_bf16_bwd; do
    nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $F -c -o build_$v/$f.o $f.cu 2> build_$v/$f.log &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsml_diag_$v.so build_$v/sml_api.o build_$v/sml_inst_f32_fwd.o build_$v/sml_inst_f32_bwd.o build_$v/sml_inst_bf16_fwd.o build_$v/sml_inst_bf16_bwd.o \
     build/sml_inst_tc.o build/sml_inst_ext_f32_fwd.o build/sml_inst_ext_f32_bwd.o build/sml_inst_ext_bf16_fwd.o build/sml_inst_ext_bf16_bwd.o \
     build/sml_inst_split_f32_fwd.o build/sml_inst_split_f32_bwd.o build/sml_inst_split_bf16_fwd.o build/sml_inst_split_bf16_bwd.o
  rm -rf build_$v
done
ls -la ../libsml_diag_*.so
