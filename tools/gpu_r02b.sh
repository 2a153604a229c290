#!/bin/bash
# round 2, call b: tests with new fixtures + full bench line (modes, parity, baselines) + ring-depth A/B + ncu launch list
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r02b_pytest_gpu.log; grep "parity " gpurun_out/r02b_pytest_gpu.log | cut -c1-400
timeout 900 python bench.py --verbose > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02b_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "sustained", d["sustained"]["value"], "e2e", d["e2e"]["value"])
print("roofline", {k: d["roofline"][k] for k in ("achieved","frac","gemm_ms_per_step","kernel_ms_per_step","whole_step_frac","sustained_whole_step_frac","other_kernels_ms_per_step")})
print("modes", d["modes"])
print("parity", {k: (v["max_abs_dlogit"], v["ordered_top5_identity_rate"], v["top5_set_identity_rate"]) for k, v in d["parity"]["modes"].items()})
print("gpu_baseline", d.get("gpu_baseline")); print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["clips_per_s_by_batch"])
print("clocks", d["clocks"], d["sustained"]["clocks"])
PY
timeout 600 python bench.py --precision fp16 --no-modes --no-staged --no-cpu-baseline --verbose > gpurun_out/r02b_bench_fp16.json 2> gpurun_out/r02b_bench_fp16.err; echo "bench fp16 rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02b_bench_fp16.json"))
print("fp16 value", d["value"], "sustained", d["sustained"]["value"], d["roofline"]["other_kernels_ms_per_step"], d["roofline"]["gemm_ms_per_step"])
PY
# ring depth A/B on the shapes of the forward
for lib in "" afft_b200/_lib/variants/libafft_s5.so afft_b200/_lib/variants/libafft_s4.so; do
  for shp in "23040 1024 1024 res" "23040 3072 1024 none" "23040 4096 1024 gelu" "23040 1024 4096 res" "4608 2048 2048 res" "4608 6144 2048 none" "4608 2048 8192 res"; do
    AFFT_B200_LIB=$lib timeout 120 python tools/gemm_time.py $shp 2>&1 | tail -1
  done
done > gpurun_out/r02b_ring_depth.txt 2>&1
cat gpurun_out/r02b_ring_depth.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 1 --no-modes --no-staged --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1; echo "ncu rc=$?"
