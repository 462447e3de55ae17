mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
rm -f gpurun_out/probe_widths.log
for F in 2 4 8; do
  FDB_NO_ROWS_KERNEL=1 timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch $F --nf $F --f0 150 >> gpurun_out/probe_widths.log 2>&1
  timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch $F --nf $F --f0 150 >> gpurun_out/probe_widths.log 2>&1
done; cat gpurun_out/probe_widths.log
timeout 400 python bench.py --stride 15 --steps 2 --no-e2e --no-cpu > gpurun_out/bench_stride15.json 2> gpurun_out/bench_stride15.err; tail -c 600 gpurun_out/bench_stride15.json
timeout 900 python bench.py --no-cpu > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_r01.json
