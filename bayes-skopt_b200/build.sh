#!/bin/sh
# Builds libbgp.so (sm_100a only) next to the Python package.  Usage: sh build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
mkdir -p build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
pids=""
for f in bgp_api bgp_gram bgp_chol bgp_small bgp_sweep bgp_acq bgp_mcmc bgp_extract bgp_grad bgp_big bgp_nccl bgp_post; do
  [ -f csrc/$f.cu ] || continue
  if [ ! -f build/$f.o ] || [ csrc/$f.cu -nt build/$f.o ] || [ csrc/bgp_common.cuh -nt build/$f.o ] \
     || [ csrc/bgp_internal.h -nt build/$f.o ] || [ ../include/bgp.h -nt build/$f.o ]; then
    nvcc $FLAGS "$@" -c csrc/$f.cu -o build/$f.o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libbgp.so build/*.o -lcudart -ldl
