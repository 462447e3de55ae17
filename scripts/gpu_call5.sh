mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
EXP_TILES=128 EXP_LPRS=2 EXP_DBGS=0,8,0,8 timeout 300 python scripts/blob_experiments.py > gpurun_out/blob_exp2.log 2>&1; cat gpurun_out/blob_exp2.log
for F in 1 2 4 8 16; do timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch $F --nf $F --f0 150 >> gpurun_out/probe_widths.log 2>&1; done; cat gpurun_out/probe_widths.log
timeout 120 python scripts/sweep_probe.py --batch 64 > gpurun_out/probe_small.log 2>&1; cat gpurun_out/probe_small.log
timeout 400 python bench.py --stride 15 --steps 2 --no-e2e --no-cpu > gpurun_out/bench_stride15.json 2> gpurun_out/bench_stride15.err; tail -c 1500 gpurun_out/bench_stride15.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 900 --csv --log-file gpurun_out/launches_r01.csv python bench.py --stride 15 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; head -c 1200 gpurun_out/bench_r01.json
