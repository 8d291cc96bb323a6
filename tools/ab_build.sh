#!/bin/bash
# tools/ab_build.sh NAME "EXTRA_NVCC_FLAGS": builds kzg_setup_powersoftau_b200/libptau_b200_NAME.so for A/B runs
# (select it with PTAU_LIB=.../libptau_b200_NAME.so)
set -e
cd "$(dirname "$0")/../kzg_setup_powersoftau_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DPTAU_FQ2_LEAF $2"
nvcc $F -c -o /tmp/k_$1.o kernels.cu &
nvcc $F -c -o /tmp/z_$1.o kzg.cu &
nvcc $F -c -o /tmp/c_$1.o capi.cu &
nvcc $F -c -o /tmp/f_$1.o files.cu &
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libptau_b200_$1.so /tmp/k_$1.o /tmp/z_$1.o /tmp/c_$1.o /tmp/f_$1.o
echo built libptau_b200_$1.so
