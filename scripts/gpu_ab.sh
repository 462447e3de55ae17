#!/bin/bash
# A/B on ONE box (run under gpurun): another revision's library (scripts/build_ab_lib.sh) against the current build, and
# the current build's own switches.  Usage: bash scripts/gpu_ab.sh [spmv|iteration|kernels|precond]
set -u
mkdir -p gpurun_out; L=gpurun_out/ab.log; rm -f $L
AB=$PWD/femder_b200/_ab/libfemder_b200_ab.so
case "${1:-spmv}" in
  spmv)        # K5 alone, 16 real slots and 8 complex slots at 1M DOF
    for i in 1 2; do
      FDB_LIB_PATH=$AB python scripts/spmv_only.py --batch 16 --real 1 --reps 30 2>&1 | sed 's/^/other   /' >> $L
      python scripts/spmv_only.py --batch 16 --real 1 --reps 30 2>&1 | sed 's/^/current /' >> $L
    done
    FDB_LIB_PATH=$AB python scripts/spmv_only.py --batch 8 --reps 30 2>&1 | sed 's/^/other   /' >> $L
    python scripts/spmv_only.py --batch 8 --reps 30 2>&1 | sed 's/^/current /' >> $L ;;
  iteration)   # whole iterations: 16 frequencies from 300 Hz at 1M DOF, 16 real slots
    for i in 1 2 3; do
      FDB_LIB_PATH=$AB python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/other   /' >> $L
      python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/current /' >> $L
    done ;;
  kernels)     # K6a alone (k_update_bj)
    for i in 1 2; do
      FDB_LIB_PATH=$AB python scripts/iter_only.py --which 1 --reps 30 2>&1 | sed 's/^/other   /' >> $L
      python scripts/iter_only.py --which 1 --reps 30 2>&1 | sed 's/^/current /' >> $L
    done ;;
  precond)     # point Jacobi against block-Jacobi in the current build: real and complex sweeps at 1M DOF
    for i in 1 2; do
      FDB_NO_BLOCK_JACOBI=1 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/jacobi  /' >> $L
      python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/block   /' >> $L
    done
    FDB_NO_BLOCK_JACOBI=1 python scripts/sweep_probe.py --shape 99,133,74 --rigid 0 --batch 8 --nf 8 --f0 200 2>&1 | sed 's/^/jacobi  /' >> $L
    python scripts/sweep_probe.py --shape 99,133,74 --rigid 0 --batch 8 --nf 8 --f0 200 2>&1 | sed 's/^/block   /' >> $L ;;
esac
cat $L
