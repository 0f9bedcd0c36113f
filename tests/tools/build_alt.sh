#!/bin/bash
# Dev builds of the library for A/B runs on one box: tests/tools/build_alt.sh name "-DFLAG ..." -> build/alt/librsfusion_<name>.so
set -e
name=$1; flags=$2
root=$(cd "$(dirname "$0")/../.." && pwd)
out=$root/build/alt; tmp=$root/build/alt/obj_$name
mkdir -p $tmp
NV="nvcc $flags -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -I$root/include -Xcompiler -fPIC,-ffp-contract=off,-fvisibility=hidden"
for f in rsf_kernels rsf_mesh rsf_api; do $NV -c $root/ra-slam_b200/csrc/$f.cu -o $tmp/$f.o & done
for f in $root/ra-slam_b200/csrc/rsf_*.cu; do b=$(basename $f .cu); case $b in rsf_kernels|rsf_mesh|rsf_api) ;; *) $NV -c $f -o $tmp/$b.o & ;; esac; done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o $out/librsfusion_$name.so $tmp/*.o
echo built $out/librsfusion_$name.so
