#!/bin/bash
# round 2, call j (2 GPUs): DataParallel drop-in on two devices, TrainState + GradBuckets at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "dataparallel" > gpurun_out/r02j_pytest_dp.log 2>&1; echo "dp rc=$?"; tail -5 gpurun_out/r02j_pytest_dp.log | cut -c1-300
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --mode train --gpus 2 --batch 16 --steps 10 > gpurun_out/r02j_train_n2.json 2> gpurun_out/r02j_train_n2.err; echo "train n2 rc=$?"; tail -3 gpurun_out/r02j_train_n2.err | cut -c1-300
python -c "
import json; d=json.load(open('gpurun_out/r02j_train_n2.json')); print('train n2', d['value'], d['ms_per_step'], d['cuda_graph'], d['config']['allreduce'])"
