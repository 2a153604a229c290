#!/bin/bash
# round 2, call n: LayerNorm / attention kernels walking their rows last-to-first (L2 recency) - A/B, plus forward tests
mkdir -p gpurun_out
for rev in 1 0 1 0; do
  AFFT_SIMT_REVERSE=$rev timeout 600 python bench.py --precision bf16 --no-staged --no-cpu-baseline --no-modes > gpurun_out/r02n_bench_rev$rev.json 2> gpurun_out/r02n_bench_rev$rev.err; echo "bench rev=$rev rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r02n_bench_rev$rev.json')); print('rev=$rev', d['value'], d['sustained']['value'], d['roofline']['gemm_ms_per_step'], d['roofline']['other_kernels_ms_per_step'])"
done
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_ops_gpu.py -m gpu -q -x -k "not large_sample" > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02n_pytest.log
