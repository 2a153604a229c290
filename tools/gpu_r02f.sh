#!/bin/bash
# round 2, call f (2 GPUs): seam tests on GPU 0, then the training step at N = 2 (GradBuckets: overlapped NCCL all-reduce in the graph)
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 900 python -m pytest tests/test_forward_gpu.py -m gpu -q -x -k "seam" > gpurun_out/r02f_pytest_seams.log 2>&1; echo "seams rc=$?"; tail -5 gpurun_out/r02f_pytest_seams.log
for comm in fp32 bf16; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --mode train --gpus 2 --batch 16 --steps 10 --grad-comm $comm > gpurun_out/r02f_train_n2_$comm.json 2> gpurun_out/r02f_train_n2_$comm.err; echo "train n2 $comm rc=$?"
  tail -c 900 gpurun_out/r02f_train_n2_$comm.json; tail -3 gpurun_out/r02f_train_n2_$comm.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --mode train --gpus 2 --batch 16 --steps 10 --no-graph > gpurun_out/r02f_train_n2_eager.json 2> gpurun_out/r02f_train_n2_eager.err; echo "train n2 eager rc=$?"; tail -c 400 gpurun_out/r02f_train_n2_eager.json
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --mode train --batch 16 --steps 10 > gpurun_out/r02f_train_n1.json 2>/dev/null; tail -c 300 gpurun_out/r02f_train_n1.json
