# A/B on ONE box: the library of commit bb76cc5 (femder_b200/_ab/libfemder_b200_old.so) against the current build
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
OLD=$PWD/femder_b200/_ab/libfemder_b200_old.so
for i in 1 2; do
  FDB_LIB_PATH=$OLD timeout 200 python scripts/spmv_only.py --batch 16 --real 1 --reps 30 2>&1 | sed 's/^/old  /' >> gpurun_out/ab.log
  timeout 200 python scripts/spmv_only.py --batch 16 --real 1 --reps 30 2>&1 | sed 's/^/new  /' >> gpurun_out/ab.log
done
FDB_LIB_PATH=$OLD timeout 200 python scripts/spmv_only.py --batch 8 --reps 30 2>&1 | sed 's/^/old  /' >> gpurun_out/ab.log
timeout 200 python scripts/spmv_only.py --batch 8 --reps 30 2>&1 | sed 's/^/new  /' >> gpurun_out/ab.log
cat gpurun_out/ab.log
