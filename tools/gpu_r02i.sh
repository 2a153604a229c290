#!/bin/bash
# round 2, call i (8 GPUs): training step at 1/2/4/8 GPUs (BASELINE config 5), T-SA / CA / Swin inference on 8 GPUs (configs 3, 4)
mkdir -p gpurun_out
run_train() {  # N batch comm tag
  local N=$1 B=$2 COMM=$3 TAG=$4
  if [ "$N" = "1" ]; then
    timeout 400 python bench.py --mode train --batch $B --steps 10 --grad-comm $COMM > gpurun_out/r02i_train_${TAG}.json 2> gpurun_out/r02i_train_${TAG}.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
      bench.py --mode train --gpus $N --batch $B --steps 10 --grad-comm $COMM > gpurun_out/r02i_train_${TAG}.json 2> gpurun_out/r02i_train_${TAG}.err
  fi
  echo "train $TAG rc=$?"
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02i_train_${TAG}.json')); print('TRAIN ${TAG}', d['n_gpus'], d['config']['clips_per_gpu_per_step'], d['value'], d['ms_per_step'], d['achieved_tflops'], d['cuda_graph'])
except Exception as e: print('TRAIN ${TAG} failed', e)"
}
for N in 1 2 4 8; do run_train $N 16 fp32 n${N}_b16_fp32; done
run_train 8 16 bf16 n8_b16_bf16
run_train 1 128 fp32 n1_b128_fp32
run_train 8 128 fp32 n8_b128_fp32
for cfg in ek100_tsa ek100_ca ek100_sa_swin; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 \
    bench.py --gpus 8 --config $cfg --no-staged --no-modes --no-cpu-baseline > gpurun_out/r02i_infer_${cfg}_n8.json 2> gpurun_out/r02i_infer_${cfg}_n8.err; echo "infer $cfg rc=$?"
  python -c "
import json
try:
    d=json.load(open('gpurun_out/r02i_infer_${cfg}_n8.json')); print('INFER ${cfg}', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['whole_step_frac'])
except Exception as e: print('INFER ${cfg} failed', e)"
done
