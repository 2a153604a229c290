#!/bin/bash
# round 2, call p: fused training nodes (MlpFn / AttnProjFn, afft_convert_dual_gelu): parity + step time at 16 / 128 clips
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_train_gpu.py -m gpu -q -x > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02p_pytest.log | cut -c1-300
for b in 16 128; do
  timeout 600 python bench.py --mode train --batch $b --steps 10 > gpurun_out/r02p_train_b$b.json 2> gpurun_out/r02p_train_b$b.err; echo "train b$b rc=$?"; tail -2 gpurun_out/r02p_train_b$b.err | cut -c1-300
  python - <<PY
import json
d=json.loads(open("gpurun_out/r02p_train_b$b.json").read().strip().splitlines()[-1])
print("b$b", d["value"], d["ms_per_step"], d.get("roofline",{}).get("whole_step_frac"))
PY
done
timeout 300 python tools/e2e_host_profile.py 256 > gpurun_out/r02p_host_profile.txt 2>&1; tail -12 gpurun_out/r02p_host_profile.txt
