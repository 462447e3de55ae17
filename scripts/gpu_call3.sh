mkdir -p gpurun_out
FDB_TILE_ROWS=64 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_tile64.log; cat gpurun_out/pytest_gpu_tile64.log
timeout 400 python scripts/blob_experiments.py > gpurun_out/blob_exp.log 2>&1; cat gpurun_out/blob_exp.log
timeout 400 python bench.py --stride 15 --steps 2 --no-e2e --no-cpu > gpurun_out/bench_stride15.json 2> gpurun_out/bench_stride15.err; tail -c 1500 gpurun_out/bench_stride15.json
timeout 120 python scripts/sweep_probe.py --batch 64 > gpurun_out/probe_small.log 2>&1; cat gpurun_out/probe_small.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/launches_small.csv python scripts/sweep_probe.py --batch 64 > gpurun_out/ncu_small.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; head -c 1200 gpurun_out/bench_r01.json
