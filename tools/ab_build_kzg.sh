#!/bin/bash
# tools/ab_build_kzg.sh NAME "EXTRA_NVCC_FLAGS": A/B build of the pairing / KZG10 kernels only (csrc/kzg.cu with the extra
# flags, linked with the other objects of the current `make`) -> kzg_setup_powersoftau_b200/libptau_b200_NAME.so
# (select it with PTAU_LIB=.../libptau_b200_NAME.so)
set -e
cd "$(dirname "$0")/../kzg_setup_powersoftau_b200/csrc"
F="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DPTAU_G1_DBL_INLINE -DPTAU_FQ2_LEAF -DPTAU_G2_DBL_INLINE $2"
nvcc $F -c -o /tmp/z_$1.o kzg.cu
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o ../libptau_b200_$1.so kernels.o /tmp/z_$1.o capi.o files.o
echo built libptau_b200_$1.so
