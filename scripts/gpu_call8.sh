mkdir -p gpurun_out; rm -f gpurun_out/ab4.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
FULL=$PWD/femder_b200/_ab/libfemder_b200_full.so
for i in 1 2; do
FDB_LIB_PATH=$FULL timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/full    /' >> gpurun_out/ab4.log
timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/packed  /' >> gpurun_out/ab4.log
done
cat gpurun_out/ab4.log
timeout 120 python scripts/sweep_probe.py --batch 16 --nf 128 --f0 100 > gpurun_out/probe_small16.log 2>&1; cat gpurun_out/probe_small16.log
timeout 120 python scripts/sweep_probe.py --batch 64 --nf 128 --f0 100 > gpurun_out/probe_small64.log 2>&1; cat gpurun_out/probe_small64.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_r01.json
