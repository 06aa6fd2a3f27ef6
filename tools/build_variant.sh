#!/bin/bash
# Build an A/B variant of the L=20 kernels only: tools/build_variant.sh NAME "-DFOO=1 ..."  -> variants/libgns_NAME.so
set -e
cd "$(dirname "$0")/../opf-graph-neural-solver_b200/csrc"
mkdir -p ../../variants
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr"
$NVCC $FLAGS $2 -c gns_inst_l20.cu -o /tmp/var_$1_l20.o
$NVCC $FLAGS $2 -c gns_backward.cu -o /tmp/var_$1_bwd.o
$NVCC $FLAGS $2 -c gns_plan.cu -o /tmp/var_$1_plan.o
$NVCC $FLAGS $2 -c gns_api.cu -o /tmp/var_$1_api.o
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o ../../variants/libgns_$1.so /tmp/var_$1_plan.o /tmp/var_$1_api.o gns_dispatch.o gns_inst_l10.o /tmp/var_$1_l20.o gns_inst_l64.o /tmp/var_$1_bwd.o -lcudart
echo built variants/libgns_$1.so
