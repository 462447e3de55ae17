#!/bin/bash
# Build the library of another git revision next to the current one, for A/B timing on ONE GPU box (boxes of the pool
# differ by ~12 % in memory bandwidth, so numbers from two gpurun calls do not compare):
#     bash scripts/build_ab_lib.sh <git-ref>      ->  femder_b200/_ab/libfemder_b200_ab.so   (git-ignored, travels with gpurun)
# then, on the box:  FDB_LIB_PATH=$PWD/femder_b200/_ab/libfemder_b200_ab.so python scripts/...   (see gpu_ab.sh)
set -eu
REF=${1:?git ref}
D=femder_b200/_ab/src
rm -rf femder_b200/_ab; mkdir -p $D/femder_b200/csrc $D/include
for f in $(git ls-tree --name-only $REF femder_b200/csrc/); do git show $REF:$f > $D/$f; done
git show $REF:include/femder_b200.h > $D/include/femder_b200.h
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --fmad=true \
     -o femder_b200/_ab/libfemder_b200_ab.so $D/femder_b200/csrc/pattern.cu $D/femder_b200/csrc/assemble.cu \
     $D/femder_b200/csrc/reorder.cu $D/femder_b200/csrc/sweep.cu $D/femder_b200/csrc/capi.cu
rm -rf $D
ls -la femder_b200/_ab/
