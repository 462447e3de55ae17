mkdir -p gpurun_out; rm -f gpurun_out/ab5.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
PREV=$PWD/femder_b200/_ab/libfemder_b200_full.so
for i in 1 2; do
FDB_LIB_PATH=$PREV timeout 100 python scripts/iter_only.py --which 1 --reps 30 2>&1 | sed 's/^/prev /' >> gpurun_out/ab5.log
timeout 100 python scripts/iter_only.py --which 1 --reps 30 2>&1 | sed 's/^/new  /' >> gpurun_out/ab5.log
done
cat gpurun_out/ab5.log
