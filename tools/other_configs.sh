#!/bin/bash
# bench.py over the non-headline BASELINE configs + strict mode (one JSON line each) -> profiles/rNN_other_configs.txt
for c in ek100_sa_tsn_wo_audio ek100_sa_swin ek100_tsa ek100_ca egtea_sa; do
  timeout 200 python bench.py --config $c --no-cpu-baseline --no-staged 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'config': '$c', 'clips_per_s': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']['value'], 'gemm_tflops': d['roofline']['achieved'], 'whole_step_tflops': d['roofline']['whole_step_tflops'], 'launches': d['launches_per_step'], 'B': d['config']['clips_per_gpu_per_step'], 'gflop_per_clip': d['config']['gemm_gflop_per_clip']}))"
done
timeout 200 python bench.py --strict --no-cpu-baseline --no-staged 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(json.dumps({'config': 'ek100_sa_tsn strict', 'clips_per_s': d['value'], 'ms_per_step': d['ms_per_step'], 'e2e': d['e2e']['value'], 'gemm_tflops_algorithmic': d['roofline']['achieved']}))"
