#!/bin/bash
# round 2, final validation: full GPU tests, the driver's bench line (both arms), ncu launch list + per-GEMM traffic, sanitizer
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02z_smoke.log
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r02z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02z_pytest_gpu.log | cut -c1-300
timeout 900 python bench.py --verbose > gpurun_out/r02z_bench.json 2> gpurun_out/r02z_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02z_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "sustained", d["sustained"]["value"], "e2e", d["e2e"]["value"], "launches", d["launches_per_step"])
print("roofline", {k: d["roofline"][k] for k in ("achieved","frac","gemm_ms_per_step","kernel_ms_per_step_profiled","whole_step_frac","sustained_whole_step_frac","other_kernels_ms_per_step")})
print("modes", {k: (v["value"], v["sustained_value"], v["whole_step_frac"]) for k, v in d["modes"].items()})
print("parity", {k: (v["max_abs_dlogit"], v["ordered_top5_identity_rate"], v["top5_set_identity_rate"], v["near_tie"], v["mismatch_clear"]) for k, v in d["parity"]["modes"].items()})
print("sweep", d["batch_sweep"]); print("gpu_baseline", d.get("gpu_baseline")); print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["clips_per_s_by_batch"])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02z_bench_reference.json 2> gpurun_out/r02z_bench_reference.err; echo "reference arm rc=$?"; cut -c1-400 gpurun_out/r02z_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-modes --no-staged --no-cpu-baseline > gpurun_out/r02z_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/summarize_ncu.py launches gpurun_out/r02z_launches_bench.csv > gpurun_out/r02z_launches_summary.txt 2>&1; head -20 gpurun_out/r02z_launches_summary.txt | cut -c1-140
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:gemm_bf16 -s 104 -c 52 --csv --log-file gpurun_out/r02z_gemm_metrics.csv python tools/ncu_forward.py 256 ek100_sa_tsn fp16 > gpurun_out/r02z_ncu_gemm.log 2>&1; echo "ncu gemm metrics rc=$?"
for c in egtea_sa ek100_tsa ek100_ca; do
  timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_train.py $c > gpurun_out/r02z_sanitize_$c.log 2>&1; echo "sanitizer $c rc=$?"; grep -E "ERROR SUMMARY|done" gpurun_out/r02z_sanitize_$c.log | tail -2
done
timeout 600 compute-sanitizer --tool memcheck python tools/ncu_forward.py 5 ek100_sa_tsn fp16 > gpurun_out/r02z_sanitize_fp16.log 2>&1; grep -E "ERROR SUMMARY" gpurun_out/r02z_sanitize_fp16.log | tail -1
for b in 16 128; do
  timeout 600 python bench.py --mode train --batch $b --steps 10 > gpurun_out/r02z_train_b$b.json 2> gpurun_out/r02z_train_b$b.err; echo "train b$b rc=$?"; tail -c 500 gpurun_out/r02z_train_b$b.json | cut -c1-500
done
