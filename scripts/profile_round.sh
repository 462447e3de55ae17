#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of the same command, full capture of the top kernel.
set -u
R=${1:-r01}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_$R.json
python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 900 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python scripts/spmv_only.py --batch 16 --real 1 --reps 5 > gpurun_out/spmv_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -o gpurun_out/spmv_real16_$R \
    python scripts/spmv_only.py --batch 16 --real 1 --reps 5 > gpurun_out/ncu_spmv.log 2>&1
echo "spmv real16 capture rc=$?"; cat gpurun_out/spmv_plain.log
python scripts/spmv_only.py --batch 8 --reps 5 > gpurun_out/spmv_plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -o gpurun_out/spmv_cpx8_$R \
    python scripts/spmv_only.py --batch 8 --reps 5 > gpurun_out/ncu_spmv8.log 2>&1
echo "spmv cpx8 capture rc=$?"; cat gpurun_out/spmv_plain8.log
