#!/bin/bash
# round 2, call k (8 GPUs): what limits the overlapped gradient all-reduce at N = 8, 16 clips per GPU
mkdir -p gpurun_out
run() {  # tag, extra env..., then bench args after --
  local TAG=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29700 \
      bench.py --mode train --gpus 8 --batch 16 --steps 10 $EXTRA > gpurun_out/r02k_${TAG}.json 2> gpurun_out/r02k_${TAG}.err
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02k_${TAG}.json')); print('K ${TAG}', d['value'], d['ms_per_step'], d['config']['allreduce'][:50])
except Exception as e: print('K ${TAG} failed', e)"
}
EXTRA="" run default X=1
EXTRA="--grad-buckets 4" run buckets4 X=1
EXTRA="--grad-buckets 1" run buckets1 X=1
EXTRA="" run nch8 NCCL_MAX_NCHANNELS=8
EXTRA="--grad-buckets 4" run nch8_b4 NCCL_MAX_NCHANNELS=8
EXTRA="" run nvls NCCL_ALGO=NVLS
EXTRA="--grad-buckets 4 --grad-comm bf16" run b4_bf16 X=1
grep -h "NVLS\|nvls" gpurun_out/r02k_nvls.err | head -3
