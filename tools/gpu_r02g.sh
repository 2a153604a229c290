#!/bin/bash
# round 2, call g: TrainState (direct gradients + native SGD) tests and the training bench at 1 GPU
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02g_pytest.log | cut -c1-300
for b in 16 128; do
  timeout 600 python bench.py --mode train --batch $b --steps 10 > gpurun_out/r02g_train_b$b.json 2> gpurun_out/r02g_train_b$b.err; echo "train b$b rc=$?"; tail -3 gpurun_out/r02g_train_b$b.err | cut -c1-300
  python -c "
import json; d=json.load(open('gpurun_out/r02g_train_b$b.json')); print('train', $b, d['value'], d['ms_per_step'], d['achieved_tflops'], d['cuda_graph'], d['config']['final_loss'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1400 --csv --log-file gpurun_out/r02g_train_launches_b128.csv \
   python bench.py --mode train --batch 128 --steps 2 --warmup 3 --no-graph > gpurun_out/r02g_ncu_train_b128.log 2>&1; echo "ncu rc=$?"
python tools/ncu_all_kernels.py gpurun_out/r02g_train_launches_b128.csv 40 | cut -c1-130
