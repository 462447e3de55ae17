mkdir -p gpurun_out; rm -f gpurun_out/ab3.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
FDB_NO_BLOCK_JACOBI=1 timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/jacobi  /' >> gpurun_out/ab3.log
timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/block   /' >> gpurun_out/ab3.log
FDB_NO_BLOCK_JACOBI=1 timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/jacobi  /' >> gpurun_out/ab3.log
timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/block   /' >> gpurun_out/ab3.log
cat gpurun_out/ab3.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 300 --csv --log-file gpurun_out/launches_bj.csv python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 > gpurun_out/ncu_bj.log 2>&1
timeout 900 python bench.py --no-cpu > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_r01.json
