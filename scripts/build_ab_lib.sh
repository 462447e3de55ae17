#!/bin/bash
# Build the library of another git revision next to the current one, for A/B timing on ONE GPU box (boxes of the pool
# differ by several per cent in memory bandwidth, so numbers from two gpurun calls do not compare):
#     bash scripts/build_ab_lib.sh <git-ref> [-DFDB_TUNING]   ->  femder_b200/_ab/libfemder_b200_ab.so   (git-ignored, travels with gpurun)
# then, on the box:  FDB_LIB_PATH=$PWD/femder_b200/_ab/libfemder_b200_ab.so python scripts/...   (see gpu_ab.sh)
# (The tuning build of the CURRENT tree is `FDB_BUILD_TUNING=1 python -m femder_b200.build`.)
set -eu
REF=${1:?git ref}
DEFS=${2:-}
D=femder_b200/_ab/src
rm -rf $D; mkdir -p $D/femder_b200/csrc $D/include
for f in $(git ls-tree --name-only $REF femder_b200/csrc/); do git show $REF:$f > $D/$f; done
git show $REF:include/femder_b200.h > $D/include/femder_b200.h
SRC=$(ls $D/femder_b200/csrc/*.cu | grep -v tc_probe)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC --fmad=true $DEFS -I$D/include \
     -o femder_b200/_ab/libfemder_b200_ab.so $SRC -lcublas -lcusolver -ldl
rm -rf $D
ls -la femder_b200/_ab/
