#!/bin/bash
# round 2, call a: fp16 mode bring-up - smoke, GPU tests, bench per precision
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python -m pytest tests -m gpu -q -x -s > gpurun_out/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r02a_pytest_gpu.log
grep "parity " gpurun_out/r02a_pytest_gpu.log
for p in bf16 fp16 strict; do
  timeout 600 python bench.py --precision $p --no-staged --no-cpu-baseline --verbose > gpurun_out/r02a_bench_$p.json 2> gpurun_out/r02a_bench_$p.err; echo "bench $p rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r02a_bench_$p.json"))
print("$p", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["clocks"])
PY
done
