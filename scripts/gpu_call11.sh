mkdir -p gpurun_out; rm -f gpurun_out/ab6.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
FDB_NO_BLOCK_JACOBI=1 timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --rigid 0 --batch 8 --nf 8 --f0 200 2>&1 | sed 's/^/jacobi  /' >> gpurun_out/ab6.log
timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --rigid 0 --batch 8 --nf 8 --f0 200 2>&1 | sed 's/^/block   /' >> gpurun_out/ab6.log
cat gpurun_out/ab6.log
