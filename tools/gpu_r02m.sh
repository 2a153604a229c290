#!/bin/bash
# round 2, call m: small-batch tile choice (2-CTA 256-row tiles vs 1-CTA 128-row tiles), fp16 after the GELU change
mkdir -p gpurun_out
for two in 1 0; do
  echo "== AFFT_GEMM_2CTA=$two"
  AFFT_GEMM_2CTA=$two timeout 600 python tools/batch_sweep.py ek100_sa_tsn bf16 8,16,32,48,64,96,128 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print(d['B'], d['max_ksplit'], d['ms_plain'], d['ms_graph'], d['clips_per_s_graph'], d['tflops_graph'])"
done > gpurun_out/r02m_small_batch_tiles.txt 2>&1
cat gpurun_out/r02m_small_batch_tiles.txt
timeout 600 python bench.py --precision fp16 --no-staged --no-cpu-baseline --no-modes > gpurun_out/r02m_bench_fp16.json 2> gpurun_out/r02m_bench_fp16.err; echo "bench fp16 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02m_bench_fp16.json')); print('fp16', d['value'], d['sustained']['value'], d['roofline']['gemm_ms_per_step'], d['roofline']['other_kernels_ms_per_step'])"
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k fp16 2>&1 | tail -2
