#!/bin/bash
# round 2, call h: full GPU test suite on the current tree + training bench (1 GPU) after the dual-conversion / attention-bwd changes
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02h_pytest_gpu.log | cut -c1-300
for b in 16 128; do
  timeout 600 python bench.py --mode train --batch $b --steps 10 > gpurun_out/r02h_train_b$b.json 2> gpurun_out/r02h_train_b$b.err; echo "train b$b rc=$?"; tail -2 gpurun_out/r02h_train_b$b.err | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/r02h_train_b$b.json')); print('train', $b, d['value'], d['ms_per_step'], d['achieved_tflops'], d['cuda_graph'], d['config']['final_loss'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 1000 --csv --log-file gpurun_out/r02h_train_launches_b128.csv \
   python bench.py --mode train --batch 128 --steps 2 --warmup 3 --no-graph > gpurun_out/r02h_ncu_train_b128.log 2>&1; echo "ncu rc=$?"
python tools/ncu_all_kernels.py gpurun_out/r02h_train_launches_b128.csv 30 | cut -c1-130
