#!/bin/bash
# round 2, call c: training path (all fusers, dropout), remaining GPU tests, training bench at 1 GPU
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -s > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/r02c_pytest_gpu.log | cut -c1-300; grep "worst relative" gpurun_out/r02c_pytest_gpu.log
for b in 16 128; do
  timeout 600 python bench.py --mode train --batch $b --steps 10 > gpurun_out/r02c_train_b$b.json 2> gpurun_out/r02c_train_b$b.err; echo "train b$b rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02c_train_b$b.json')); print('train', $b, d['value'], d['ms_per_step'], d['achieved_tflops'], d['cuda_graph'], d['config']['final_loss'])"
done
timeout 600 python bench.py --mode train --batch 16 --steps 10 --config ek100_tsa > gpurun_out/r02c_train_tsa.json 2> gpurun_out/r02c_train_tsa.err; echo "train tsa rc=$?"; tail -c 600 gpurun_out/r02c_train_tsa.json
timeout 600 python bench.py --mode train --batch 16 --steps 10 --config ek100_ca > gpurun_out/r02c_train_ca.json 2> gpurun_out/r02c_train_ca.err; echo "train ca rc=$?"; tail -c 600 gpurun_out/r02c_train_ca.json
