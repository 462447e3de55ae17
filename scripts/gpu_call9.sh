mkdir -p gpurun_out; rm -f gpurun_out/probe_small_full.log
for B in 16 64; do timeout 200 python scripts/sweep_probe.py --batch $B --nf 481 --f0 20 >> gpurun_out/probe_small_full.log 2>&1; done
FDB_NO_BLOCK_JACOBI=1 timeout 200 python scripts/sweep_probe.py --batch 16 --nf 481 --f0 20 2>&1 | sed 's/^/jacobi /' >> gpurun_out/probe_small_full.log
cat gpurun_out/probe_small_full.log
timeout 900 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_r01.json
