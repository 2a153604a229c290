#!/bin/bash
# round 2, call l: Identity dim_encoder / dim_decoder fixtures, fp_output_attentions, normalised roofline slices
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_ops_gpu.py -m gpu -q -x -k "identity or attentions or seam or layernorm or head_and_mapping" > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02l_pytest.log | cut -c1-300
timeout 600 python bench.py --no-modes --no-staged --no-cpu-baseline > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02l_bench.json"))
r=d["roofline"]
print("value", d["value"], "ms", d["ms_per_step"], "sustained", d["sustained"]["value"])
print({k: r[k] for k in ("achieved","frac","gemm_ms_per_step","kernel_ms_per_step_profiled","other_kernels_ms_per_step","gemm_share_of_kernel_time")})
PY
