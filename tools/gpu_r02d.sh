#!/bin/bash
# round 2, call d: TMA-staged epilogue (v2) bring-up: op tests, forward parity, A/B timing v2 on/off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x > gpurun_out/r02d_pytest_ops.log 2>&1; echo "ops rc=$?"; tail -5 gpurun_out/r02d_pytest_ops.log
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_train_gpu.py -m gpu -q -x -k "not large_sample" > gpurun_out/r02d_pytest_fwd.log 2>&1; echo "fwd rc=$?"; tail -5 gpurun_out/r02d_pytest_fwd.log
for v2 in 1 0; do
  for shp in "23040 1024 1024 res" "23040 3072 1024 none" "23040 4096 1024 gelu" "23040 1024 4096 res" "4608 2048 2048 res" "4608 6144 2048 none" "4608 2048 8192 res" "4864 3806 1024 f32"; do
    echo -n "v2=$v2 "; AFFT_GEMM_EPI_V2=$v2 timeout 120 python tools/gemm_time.py $shp 2>&1 | tail -1
  done
done > gpurun_out/r02d_epi_v2_ab.txt 2>&1
cat gpurun_out/r02d_epi_v2_ab.txt
for v2 in 1 0; do
  AFFT_GEMM_EPI_V2=$v2 timeout 600 python bench.py --no-modes --no-staged --no-cpu-baseline --verbose > gpurun_out/r02d_bench_v2_$v2.json 2> gpurun_out/r02d_bench_v2_$v2.err; echo "bench v2=$v2 rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r02d_bench_v2_$v2.json"))
print("v2=$v2 value", d["value"], "sustained", d["sustained"]["value"], "gemm_ms", d["roofline"]["gemm_ms_per_step"], d["roofline"]["frac"], d["roofline"]["other_kernels_ms_per_step"])
PY
done
