#!/bin/bash
# Run on the GPU box (under gpurun): bench line, ncu launch list of a sweep, full captures of the iteration's kernels.
set -u
R=${1:-r02}
mkdir -p gpurun_out
python scripts/sweep_for_ncu.py > gpurun_out/sweep_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "sweep/" --csv --log-file gpurun_out/launches_$R.csv \
    python scripts/sweep_for_ncu.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/sweep_plain.log
cap() {  # name, kernel regex, script args
  python scripts/iter_only.py $3 > gpurun_out/plain_$1.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/$1_$R python scripts/iter_only.py $3 > gpurun_out/ncu_$1.log 2>&1
  echo "$1 capture rc=$?"; tail -1 gpurun_out/plain_$1.log
  ncu -i gpurun_out/$1_$R.ncu-rep --page raw --csv > gpurun_out/$1_$R.raw.csv 2>/dev/null
}
cap precond_fused k_precond_fused "--which 1"
cap coarse_finish k_coarse_finish "--which 6"
cap coarse_project k_modal_project "--which 3"
cap coarse_expand k_modal_expand "--which 4"
cap spmv_real16 k_spmv_blob "--which 0 --modal 0"
cap pupdate_tiles_real16 k_pupdate_tiles "--which 2 --modal 0"
# gpurun merges at most 64 MiB back: keep the two reports worth opening with the source page, the others as raw CSV
rm -f gpurun_out/pupdate_tiles_real16_$R.ncu-rep gpurun_out/spmv_real16_$R.ncu-rep gpurun_out/coarse_project_$R.ncu-rep gpurun_out/coarse_expand_$R.ncu-rep
# the bench line last, so that its `traffic` fields come from this round's captures
python scripts/summarize_profiles.py $R > /dev/null 2>&1
python bench.py > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
echo "bench rc=$?"; head -c 300 gpurun_out/bench_$R.json; echo
ls -la gpurun_out | grep $R
