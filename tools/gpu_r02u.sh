#!/bin/bash
# round 2, call u (2 GPUs): training step with the fused nodes + MixUp at N = 1 and N = 2 on the same box
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 300 python bench.py --mode train --batch 16 --steps 10 > gpurun_out/r02u_train_n1.json 2> gpurun_out/r02u_train_n1.err; echo "n1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --mode train --gpus 2 --batch 16 --steps 10 > gpurun_out/r02u_train_n2.json 2> gpurun_out/r02u_train_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/r02u_train_n2.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --mode train --gpus 2 --batch 128 --steps 10 > gpurun_out/r02u_train_n2_b128.json 2> gpurun_out/r02u_train_n2_b128.err; echo "n2 b128 rc=$?"
python - <<'PY'
import json
for t in ("n1", "n2", "n2_b128"):
    try:
        d = json.loads(open(f"gpurun_out/r02u_train_{t}.json").read().strip().splitlines()[-1])
        print(t, d["value"], d["ms_per_step"], d["cuda_graph"], d["config"]["final_loss"], d["config"]["allreduce"][:60])
    except Exception as e:
        print(t, "failed", e)
PY
