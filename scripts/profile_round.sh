#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of the same command, full capture of the top kernel.
set -u
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r01.json
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/launches_r01.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python scripts/spmv_only.py --batch 8 --reps 5 > gpurun_out/spmv_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -o gpurun_out/spmv_r01 \
    python scripts/spmv_only.py --batch 8 --reps 5 > gpurun_out/ncu_spmv.log 2>&1
echo "spmv capture rc=$?"; cat gpurun_out/spmv_plain.log
