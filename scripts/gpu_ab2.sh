# A/B on ONE box: whole iterations (16 real slots at 1M DOF, 16 frequencies from 300 Hz), old library against the current build
mkdir -p gpurun_out; rm -f gpurun_out/ab2.log
OLD=$PWD/femder_b200/_ab/libfemder_b200_old.so
for i in 1 2 3; do
  FDB_LIB_PATH=$OLD timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/old  /' >> gpurun_out/ab2.log
  timeout 200 python scripts/sweep_probe.py --shape 99,133,74 --batch 16 --nf 16 --f0 300 2>&1 | sed 's/^/new  /' >> gpurun_out/ab2.log
done
cat gpurun_out/ab2.log
