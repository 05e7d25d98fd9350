#!/bin/bash
# Builds nudge_b200/lib/libnudge_b200.so for sm_100a.  -fmad=false: FMAs only where the reference has madd/msub
# (see nb_common.cuh); IEEE division/sqrt; no flush-to-zero.
set -e
cd "$(dirname "$0")"
mkdir -p ../lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC -Xcompiler -O2"
g++ -O2 -fPIC -msse2 -c nb_lut_host.cpp -o ../lib/nb_lut_host.o
$NVCC $FLAGS ${NB_PTXAS_V:+-Xptxas -v} -c nb_api.cu -o ../lib/nb_api.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libnudge_b200.so ../lib/nb_api.o ../lib/nb_lut_host.o -lcudart_static -ldl -lrt -lpthread
g++ -O2 -fPIC -shared -o ../lib/libnudge_compat.so nudge_compat.cpp -L../lib -lnudge_b200 -Wl,-rpath,'$ORIGIN'
mkdir -p ../../tools/bin
g++ -O2 -std=c++17 -o ../../tools/bin/nb_replay ../../tools/nb_replay.cpp -L../lib -lnudge_b200 -Wl,-rpath,'$ORIGIN/../../nudge_b200/lib'
echo "built nudge_b200/lib/libnudge_b200.so and libnudge_compat.so"
