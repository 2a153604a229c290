#!/bin/bash
# round 2, call t: training iteration with MixUp + soft-label losses (afft_b200/runner.py): parity test, step time with / without
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -x -k "mixup" > gpurun_out/r02t_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02t_pytest.log | cut -c1-300
for spec in "16 " "128 " "16 --no-mixup" "16 --no-graph"; do
  set -- $spec
  tag="b$1$2"
  timeout 600 python bench.py --mode train --batch $1 $2 --steps 10 > gpurun_out/r02t_train_$tag.json 2> gpurun_out/r02t_train_$tag.err; echo "train $tag rc=$?"; tail -2 gpurun_out/r02t_train_$tag.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02t_train_$tag.json").read().strip().splitlines()[-1])
print("$tag", d["value"], d["ms_per_step"], d["cuda_graph"], d["config"]["final_loss"], d["config"]["mixup"][:40])
PY
done
