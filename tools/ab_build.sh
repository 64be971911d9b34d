#!/bin/sh
# A/B library for kernel experiments: tools/_ab/libccsdt_b200_<name>.so built with extra -D flags.
#   usage: tools/ab_build.sh <name> [-DFLAG ...]      then  CCSDT_B200_LIB=tools/_ab/libccsdt_b200_<name>.so python ...
set -e
cd "$(dirname "$0")/../exachem_b200/csrc"
name=$1; shift
out=../../tools/_ab/libccsdt_b200_$name.so
mkdir -p ../../tools/_ab/_obj_$name
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ccbin g++"
for f in ccsdt_kernels ccsdt_capi ccsdt_store ccsdt_comm ccsdt_share ccsdt_v2; do
  $NV "$@" -c $f.cu -o ../../tools/_ab/_obj_$name/$f.o
done
g++ -std=c++17 -O2 -fPIC -fvisibility=hidden -c ccsdt_host.cpp -o ../../tools/_ab/_obj_$name/ccsdt_host.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out ../../tools/_ab/_obj_$name/*.o -Xcompiler -fvisibility=hidden -ldl -lrt -lpthread
echo built $out
