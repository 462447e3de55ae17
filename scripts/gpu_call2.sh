mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python scripts/blob_experiments.py > gpurun_out/blob_exp.log 2>&1; cat gpurun_out/blob_exp.log
timeout 400 python bench.py --stride 15 --steps 2 --no-e2e --no-cpu > gpurun_out/bench_stride15.json 2> gpurun_out/bench_stride15.err; tail -c 1500 gpurun_out/bench_stride15.json
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; head -c 1200 gpurun_out/bench_r01.json
