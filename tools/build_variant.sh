#!/bin/bash
# builds an instrumented copy of libsgp.so for timing experiments: tools/build_variant.sh <name> <extra nvcc flags...>
# -> gaussianprocessnode_b200/libsgp_<name>.so ; use it with SGP_LIB_PATH=...
set -e
cd "$(dirname "$0")/../gaussianprocessnode_b200"
name=$1; shift
for f in api sweep sweep_se sweep_m32 sweep_m52 sweep4_se sweep4_m32 sweep4_m52 dense dense_coop uncertain theta inmsg comm; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c csrc/$f.cu -o build/${f}_$name.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libsgp_$name.so build/api_$name.o build/sweep_$name.o build/sweep_se_$name.o build/sweep_m32_$name.o build/sweep_m52_$name.o build/sweep4_se_$name.o build/sweep4_m32_$name.o build/sweep4_m52_$name.o build/dense_$name.o build/dense_coop_$name.o build/uncertain_$name.o build/theta_$name.o build/inmsg_$name.o build/comm_$name.o -lcudart -ldl
echo libsgp_$name.so
