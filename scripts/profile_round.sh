#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of the same solver, full captures of the iteration's kernels.
set -u
R=${1:-r01}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "bench rc=$?"; head -c 400 gpurun_out/bench_$R.json; echo
python bench.py --stride 15 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/bench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 900 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --stride 15 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
cap() {  # name, kernel regex, script args
  python scripts/iter_only.py $3 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$1_$R python scripts/iter_only.py $3 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 capture rc=$?"
  ncu -i gpurun_out/$1_$R.ncu-rep --page raw --csv > gpurun_out/$1_$R.raw.csv 2>/dev/null
}
cap update_bj_real16 k_update_bj "--which 1"
cap spmv_real16 k_spmv_blob "--which 0"
cap pupdate_bj_real16 k_pupdate_bj "--which 2"
python scripts/spmv_only.py --batch 8 --reps 5 > gpurun_out/spmv_plain8.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_spmv -s 3 -c 1 -f -o gpurun_out/spmv_cpx8_$R \
    python scripts/spmv_only.py --batch 8 --reps 5 > gpurun_out/ncu_spmv8.log 2>&1
echo "spmv cpx8 capture rc=$?"; cat gpurun_out/spmv_plain8.log
ncu -i gpurun_out/spmv_cpx8_$R.ncu-rep --page raw --csv > gpurun_out/spmv_cpx8_$R.raw.csv 2>/dev/null
# gpurun merges at most 64 MiB back: keep the two reports worth opening with the source page, the others as raw CSV
rm -f gpurun_out/pupdate_bj_real16_$R.ncu-rep gpurun_out/spmv_cpx8_$R.ncu-rep
ls -la gpurun_out | head -40
