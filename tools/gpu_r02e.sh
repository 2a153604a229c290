#!/bin/bash
# round 2, call e: launch list of one eager training step (B = 16 and B = 128), full GPU test run, default bench line
mkdir -p gpurun_out
for b in 16 128; do
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 6000 -c 2500 --csv --log-file gpurun_out/r02e_train_launches_b$b.csv \
     python bench.py --mode train --batch $b --steps 2 --warmup 3 --no-graph > gpurun_out/r02e_ncu_train_b$b.log 2>&1; echo "ncu train b$b rc=$?"
  python tools/ncu_all_kernels.py gpurun_out/r02e_train_launches_b$b.csv 45 > gpurun_out/r02e_train_kernels_b$b.txt 2>&1
  head -50 gpurun_out/r02e_train_kernels_b$b.txt
done
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02e_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02e_pytest_gpu.log
